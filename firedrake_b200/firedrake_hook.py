"""Hook for a REAL Firedrake installation.

**Untestable in this image** -- Firedrake, PyOP2, UFL and PETSc are not
importable here, so nothing below has ever run; it is
the reference-side binding a maintainer would start from, kept next to the
engine so that the call sequence it needs stays in sync with
``include/fdb200.h``.  The same sequence IS exercised, end to end, by
``firedrake_b200.op2`` (the PyOP2 mirror) in ``tests/``.

Usage (where Firedrake exists)::

    import firedrake_b200.firedrake_hook as hook
    hook.install()            # before the first assemble()
    ...
    hook.uninstall()

What it does: replaces ``pyop2.global_kernel.compile_global_kernel``
(reference pyop2/global_kernel.py:426-456).  For global kernels whose local
kernel was generated from one of the supported forms it returns a callable with
the JIT-compiled wrapper's signature ``fn(start, end, *arglist)`` that forwards
to ``fdb_kernel_call`` in host-pointer mode; everything else falls through to
the stock C path.  A form is recognised by comparing its UFL signature with
template forms built on the same function space -- the reference keys its own
kernel cache on ``form.signature()`` (firedrake/tsfc_interface.py:55-62).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_original = None
_registry = {}          # form signature -> dict(alpha, beta, degree, cdim, rank)


def register_form(form, alpha, beta):
    """Declare that UFL ``form`` is ``alpha*inner(grad u, grad v)*dx +
    beta*inner(u, v)*dx`` on a Q_p (x) P_p space with ``dx(degree=2p)``.

    Called by user code (or by :func:`install` for the demos' forms) because
    recognising the algebra of an arbitrary UFL form is the form compiler's job;
    here only forms that were explicitly registered are offloaded."""
    args = form.arguments()
    V = args[0].function_space()
    el = V.ufl_element()
    degree = el.degree()
    if isinstance(degree, tuple):
        degree = degree[0]
    _registry[form.signature()] = dict(alpha=float(alpha), beta=float(beta), degree=int(degree),
                                       cdim=int(V.value_size), rank=len(args))


def _descriptor_for(global_kernel):
    """KernelDesc for a pyop2 GlobalKernel generated from a registered form,
    else None.  The TSFC kernel carries its form signature in the cache key
    attached by firedrake.tsfc_interface (KernelInfo / TSFCKernel)."""
    lk = global_kernel.local_kernel
    sig = getattr(lk, "form_signature", None)     # attached by a one-line patch in tsfc_interface
    meta = _registry.get(sig)
    if meta is None:
        return None
    from .fiat_lite import interval_element
    el = interval_element(meta["degree"])
    n = meta["degree"] + 1
    d = _lib.KernelDesc()
    d.form, d.rank = _lib.FORM_HELMHOLTZ, meta["rank"]
    d.cell = _lib.CELL_HEX_EXTRUDED if global_kernel._extruded else _lib.CELL_HEX
    d.integral, d.degree, d.nq, d.cdim = _lib.INTEGRAL_CELL, meta["degree"], n, meta["cdim"]
    d.scatter = _lib.SCATTER_ATOMIC
    d.alpha, d.beta = meta["alpha"], meta["beta"]
    for q in range(n):
        d.wq[q], d.xq[q] = el.wq[q], el.xq[q]
        for a in range(n):
            d.B[q * n + a], d.D[q * n + a] = el.B[q, a], el.D[q, a]
    # Map.offset of the argument map and of the coordinate map are compile-time
    # constants of the wrapper (pyop2/codegen/builder.py:52-58)
    maps = [m for arg in global_kernel.arguments for m in getattr(arg, "maps", ())]
    keep = []
    if global_kernel._extruded:
        o0 = np.ascontiguousarray(maps[0].offset, dtype=np.int32)
        o1 = np.ascontiguousarray(maps[1].offset, dtype=np.int32)
        keep = [o0, o1]
        d.offset0 = o0.ctypes.data_as(C.POINTER(C.c_int32))
        d.offset1 = o1.ctypes.data_as(C.POINTER(C.c_int32))
    return d, keep


def _make_wrapper(handle, global_kernel):
    L = _lib.lib()
    extruded = global_kernel._extruded
    subset = global_kernel._subset

    def fn(start, end, *arglist):
        """Same positional protocol as the generated ``wrap_<kernel>``:
        [layers] [subset] data pointers (TSFC order) ... map pointers."""
        pos = 0
        layers = subset_ptr = None
        if extruded:
            layers, pos = arglist[pos], pos + 1
        if subset:
            subset_ptr, pos = arglist[pos], pos + 1
        nargs = 3 if global_kernel.local_kernel.num_args == 3 else 2
        data = arglist[pos:pos + nargs]
        maps = arglist[pos + nargs:]
        ca = _lib.CallArgs()
        ca.start, ca.end = start, end
        ca.layers = C.cast(layers, C.POINTER(C.c_int32)) if layers else None
        ca.subset = subset_ptr
        ca.nargs, ca.args = nargs, (C.c_void_p * nargs)(*data)
        # sizes and dat_versions are not part of the reference arglist: the patched
        # Parloop passes them through a side channel (Parloop._fdb_sizes / _fdb_versions)
        ca.arg_bytes = (C.c_size_t * nargs)(*fn.sizes[:nargs])
        ca.arg_versions = (C.c_uint64 * nargs)(*fn.versions[:nargs])
        ca.nmaps, ca.maps = len(maps), (C.c_void_p * len(maps))(*maps)
        ca.map_bytes = (C.c_size_t * len(maps))(*fn.map_sizes)
        ca.location, ca.writeback, ca.output_is_zero = _lib.LOC_HOST, 1, int(fn.output_is_zero)
        _lib.check(L.fdb_kernel_call(handle, C.byref(ca)), "fdb_kernel_call")
        return 0

    fn.sizes, fn.versions, fn.map_sizes, fn.output_is_zero = (), (), (), False
    return fn


def install():
    """Swap ``compile_global_kernel``; idempotent."""
    global _original
    import pyop2.global_kernel as gk          # noqa: F401  (ImportError here: no Firedrake)
    if _original is not None:
        return
    _original = gk.compile_global_kernel

    def compile_global_kernel(kernel, comm):
        found = _descriptor_for(kernel)
        if found is None:
            return _original(kernel, comm)
        desc, keep = found
        L = _lib.init()
        h = C.c_void_p()
        _lib.check(L.fdb_kernel_create(C.byref(desc), C.byref(h)), "fdb_kernel_create")
        del keep
        return _make_wrapper(h, kernel)

    gk.compile_global_kernel = compile_global_kernel


def uninstall():
    global _original
    if _original is not None:
        import pyop2.global_kernel as gk
        gk.compile_global_kernel = _original
        _original = None
