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
(reference pyop2/global_kernel.py:426-456) with a three-way choice:

1. FAST PATH -- global kernels whose local kernel was generated from a
   registered form of the supported set get a callable with the JIT-compiled
   wrapper's signature ``fn(start, end, *arglist)`` that forwards to the
   hand-written sm_100a kernels (``fdb_kernel_create`` descriptor).  A form is
   recognised by its UFL signature -- the reference keys its own kernel cache on
   ``form.signature()`` (firedrake/tsfc_interface.py:55-62).
2. GENERIC PATH -- any other global kernel whose arguments are Dats and Globals:
   the local kernel's C source (a ``CStringLocalKernel``'s string, or
   ``loopy.generate_code_v2`` of a ``LoopyLocalKernel``, which is what PyOP2
   itself inlines into its wrapper) is handed to the engine's NVRTC wrapper
   builder (``fdb_wrapper_create``), described by the same ``*KernelArg``
   objects PyOP2's ``WrapperBuilder`` consumes (pyop2/global_kernel.py:26-170,
   pyop2/codegen/builder.py:840-916).
3. Everything else (PETSc ``Mat`` arguments -- the engine assembles into its own
   CSR, not into a PETSc handle --, MixedDats, periodic extrusion, variable
   layers) falls through to the stock C path.
"""
from __future__ import annotations

import ctypes as C
import itertools
import weakref

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


_NP_DTYPES = {"float64": _lib.F64, "float32": _lib.F32, "int32": _lib.I32, "uint32": _lib.U32,
              "int64": _lib.I64}


def _local_kernel_source(lk):
    """C source of a pyop2 LocalKernel (pyop2/local_kernel.py:186-260)."""
    if isinstance(lk.code, str):
        return lk.code
    import loopy as lp
    return lp.generate_code_v2(lk.code).device_code()


def _generic_for(global_kernel):
    """(WrapperDesc, keepalive) for a pyop2 GlobalKernel made of Dat / Global arguments,
    else None.  Mirrors WrapperBuilder.add_argument (pyop2/codegen/builder.py:840-916)."""
    import pyop2.global_kernel as gk
    from pyop2.types import IterationRegion
    g = global_kernel
    if g._extruded and (not g._constant_layers or g._extruded_periodic):
        return None
    region = {None: _lib.REGION_ALL, IterationRegion.ALL: _lib.REGION_ALL,
              IterationRegion.BOTTOM: _lib.REGION_ON_BOTTOM, IterationRegion.TOP: _lib.REGION_ON_TOP,
              IterationRegion.INTERIOR_FACETS: _lib.REGION_ON_INTERIOR_FACETS}[g._iteration_region]
    horizontal = region == _lib.REGION_ON_INTERIOR_FACETS
    lk = g.local_kernel
    slots, keep = [], []          # distinct MapKernelArgs by identity, first-use order

    def slot(m):
        base = m.base_map if isinstance(m, gk.PermutedMapKernelArg) else m
        for i, b in enumerate(slots):
            if b is base:
                return i
        slots.append(base)
        return len(slots) - 1

    def ints(v):
        if v is None:
            return None
        a = np.ascontiguousarray(v, dtype=np.int32)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int32))

    arr = (_lib.WrapperArg * len(g.arguments))()
    for w, larg, garg in zip(arr, lk.arguments, g.arguments):
        w.access = int(larg.access.value)
        w.dtype = _NP_DTYPES.get(np.dtype(larg.dtype).name, 0)
        w.map = w.map2 = -1
        if w.dtype == 0:
            return None
        if isinstance(garg, gk.GlobalKernelArg) and not garg.double:
            w.kind, w.dim = _lib.ARG_GLOBAL, int(np.prod(garg.dim, dtype=int))
        elif isinstance(garg, gk.DatKernelArg) and garg.index is None:
            w.kind, w.dim = _lib.ARG_DAT, int(np.prod(garg.dim, dtype=int))
            m = garg.map_
            if isinstance(m, gk.ComposedMapKernelArg):
                return None
            if m is not None:
                base = m.base_map if isinstance(m, gk.PermutedMapKernelArg) else m
                if base.offset_quotient is not None:
                    return None
                w.map, w.arity = slot(m), base.arity
                perm = m.permutation if isinstance(m, gk.PermutedMapKernelArg) else None
                w.permutation = ints(perm)
                off = base.offset
                if off is not None and perm is not None:
                    off = np.asarray(off)[np.asarray(perm)]
                w.offset = ints(off)
                w.interior_horizontal = int(horizontal)
        else:
            return None               # Mat / MixedDat / passthrough: stock path
    d = _lib.WrapperDesc()
    src = _local_kernel_source(lk).encode()
    name = lk.name.encode()
    keep += [arr, src, name]
    d.kernel_source, d.kernel_name = src, name
    d.nargs, d.args = len(g.arguments), arr
    d.extruded, d.subset, d.iteration_region = int(g._extruded), int(g._subset), region
    d.pass_layer_arg = int(bool(g._pass_layer_arg))
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
        nargs = len(global_kernel.arguments)
        data = arglist[pos:pos + nargs]
        maps = arglist[pos + nargs:]
        ca = _lib.CallArgs()
        ca.start, ca.end = start, end
        ca.layers = C.cast(layers, C.POINTER(C.c_int32)) if layers else None
        ca.subset = subset_ptr
        ca.nargs, ca.args = nargs, (C.c_void_p * nargs)(*data)
        # sizes and dat_versions are not part of the reference arglist: _patched_compute
        # (below) sets them on this callable right before PyOP2 invokes it
        ca.arg_bytes = (C.c_size_t * nargs)(*fn.sizes[:nargs])
        if fn.versions is not None:        # None: no sound cache key -> every call re-uploads
            ca.arg_versions = (C.c_uint64 * nargs)(*fn.versions[:nargs])
        ca.nmaps, ca.maps = len(maps), (C.c_void_p * len(maps))(*maps)
        ca.map_bytes = (C.c_size_t * len(maps))(*fn.map_sizes)
        if fn.map_generations:
            ca.map_versions = (C.c_uint64 * len(maps))(*fn.map_generations)
        ca.location, ca.writeback, ca.output_is_zero = _lib.LOC_HOST, 1, int(fn.output_is_zero)
        _lib.check(L.fdb_kernel_call(handle, C.byref(ca)), "fdb_kernel_call")
        return 0

    fn.sizes, fn.versions, fn.map_sizes, fn.output_is_zero, fn.map_generations = (), (), (), False, ()
    return fn


_generations = itertools.count(1)


def _drop_mirror(ptr):
    try:
        if _lib._initialised is not None:
            _lib._lib.fdb_mirror_drop(ptr)
    except Exception:
        pass


def _track(obj, buf):
    """First sight of a PyOP2 carrier whose host buffer the engine mirrors: give it a
    generation id (never reused, so a new object at a recycled address misses the cache) and
    release its mirror when it is garbage-collected (no unbounded growth in time loops; the
    engine additionally bounds the cache by LRU eviction, runtime.cu)."""
    gen = getattr(obj, "_fdb_generation", None)
    if gen is None:
        gen = next(_generations)
        try:
            obj._fdb_generation = gen
            if buf is not None:
                weakref.finalize(obj, _drop_mirror, buf.ctypes.data)
        except (AttributeError, TypeError):
            return 0                      # no attribute slot: address-keyed only
    return gen


_original_compute = None


def _patched_compute(self, part):
    """``Parloop._compute`` (pyop2/parloop.py:224-232) with the side channel the engine's
    host-pointer mode needs and the reference arglist lacks: byte sizes of the argument and map
    buffers and ``dat_version`` of every argument (the mirror cache key).  PyOP2 bumps the version
    of written arguments BEFORE the compute phases (``increment_dat_version``,
    parloop.py:243-245) while the engine records ``version + 1`` after its write-back, so written
    arguments are announced with ``dat_version - 1``."""
    import pyop2.global_kernel as gk
    fn = gk.compile_global_kernel(self.global_kernel, self.comm)      # memory-cached
    if hasattr(fn, "sizes"):
        from pyop2.types import READ
        sizes, versions = [], []
        # (buffer address, dat_version) is a sound mirror key only while nothing rewrites the
        # buffer in place without bumping the version.  Under MPI the PetscSF halo exchange
        # rewrites ghost rows of READ Dats between the core and owned phases, and INC ghost
        # zeroing does the same: no versions there, every call re-uploads.
        cacheable = self.comm.size == 1
        for arg, access in zip(self.arguments, self.accesses):
            data = arg.data
            for d in (data if hasattr(data, "__iter__") and not hasattr(data, "_data") else (data,)):
                buf = getattr(d, "_data", None)
                sizes.append(0 if buf is None else buf.nbytes)
                v = getattr(d, "dat_version", 0)
                versions.append(v if access is READ else max(v - 1, 0))
                _track(d, buf)
                if getattr(getattr(d, "dataset", None), "halo", None) is not None and self.comm.size > 1:
                    cacheable = False
        maps = {m: None for d in self.arguments for m in d.map_kernel_args}
        fn.sizes, fn.versions = tuple(sizes), (tuple(versions) if cacheable else None)
        fn.map_generations = tuple(_track(m, getattr(m, "values_with_halo", None)) for m in maps)
        fn.map_sizes = tuple(self.iterset.total_size * 4 * getattr(m, "arity", 0) if not hasattr(m, "nbytes")
                             else m.nbytes for m in maps)
        fn.output_is_zero = False
    return _original_compute(self, part)


def install():
    """Swap ``compile_global_kernel`` and ``Parloop._compute``; idempotent."""
    global _original, _original_compute
    import pyop2.global_kernel as gk          # noqa: F401  (ImportError here: no Firedrake)
    import pyop2.parloop as pl
    if _original is not None:
        return
    _original = gk.compile_global_kernel
    _original_compute = pl.Parloop._compute
    pl.Parloop._compute = _patched_compute

    def compile_global_kernel(kernel, comm):
        found = _descriptor_for(kernel)
        h = C.c_void_p()
        if found is not None:                      # 1. hand-written fast path
            desc, keep = found
            _lib.check(_lib.init().fdb_kernel_create(C.byref(desc), C.byref(h)), "fdb_kernel_create")
            del keep
            return _make_wrapper(h, kernel)
        found = _generic_for(kernel)
        if found is not None:                      # 2. NVRTC wrapper around the local kernel's C
            desc, keep = found
            _lib.check(_lib.init().fdb_wrapper_create(C.byref(desc), C.byref(h)), "fdb_wrapper_create")
            del keep
            return _make_wrapper(h, kernel)
        return _original(kernel, comm)             # 3. stock C path

    gk.compile_global_kernel = compile_global_kernel


def uninstall():
    global _original, _original_compute
    if _original is not None:
        import pyop2.global_kernel as gk
        import pyop2.parloop as pl
        gk.compile_global_kernel = _original
        pl.Parloop._compute = _original_compute
        _original = _original_compute = None
