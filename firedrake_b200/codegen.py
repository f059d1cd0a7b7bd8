"""Generic parloops: arbitrary C local kernels behind ``op2.par_loop``.

Mirrors the reference's C-string kernel route -- ``op2.Kernel(code, name)`` ->
``CStringLocalKernel`` (pyop2/local_kernel.py:186-207) -> ``WrapperBuilder``
(pyop2/codegen/builder.py:702-1008) -> host compiler -- with the engine's
NVRTC wrapper builder (``fdb_wrapper_*`` in include/fdb200.h,
csrc/wrapper_jit.cu): the description of the parloop's arguments is turned into
a ``fdb_wrapper_desc``, the engine generates the sm_100a global kernel around the
local kernel source and compiles it at run time.

The argument description (access, dtype, dims, map slots, offsets,
permutations) is exactly what PyOP2 folds into ``GlobalKernel.cache_key``
(pyop2/global_kernel.py:309-317), so it is also the cache key here.

Status: the generated code is checked on the CPU (tests/test_codegen.py: NVRTC
compile for sm_100a, and a host re-compilation of the generated wrapper body run
against the reference's golden arrays) and on the GPU (tests/test_jit_gpu.py, green on a
B200 since round 2).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

_DTYPE_CODE = {
    np.dtype(np.float64): _lib.F64, np.dtype(np.float32): _lib.F32,
    np.dtype(np.int32): _lib.I32, np.dtype(np.uint32): _lib.U32, np.dtype(np.int64): _lib.I64,
}
_REGIONS = {"ALL": _lib.REGION_ALL, "ON_BOTTOM": _lib.REGION_ON_BOTTOM, "ON_TOP": _lib.REGION_ON_TOP,
            "ON_INTERIOR_FACETS": _lib.REGION_ON_INTERIOR_FACETS}


@dataclass(frozen=True)
class CStringKernel:
    """``op2.Kernel(code, name)``: a local kernel given as C source
    (pyop2/local_kernel.py:33-43, 186-207).  Every argument is a flat pointer to
    the packed local data of one parloop argument, in parloop argument order."""
    code: str
    name: str
    accesses: tuple = None
    flop_count: int = 0
    opts: dict = field(default=None, compare=False, hash=False)

    def __post_init__(self):
        if not isinstance(self.name, str) or not self.name.isidentifier():
            raise ValueError("kernel name must be a C identifier")


class PermutedMap:
    """``op2.PermutedMap(map, permutation)``: ``local[i] = global[map[permutation[i]]]``
    (pyop2/types/map.py:172-230).  Adds no new map argument to the wrapper."""

    def __init__(self, map_, permutation):
        self.map_ = map_
        self.permutation = np.ascontiguousarray(permutation, dtype=np.int32)
        if sorted(self.permutation.tolist()) != list(range(map_.arity)):
            raise ValueError("permutation must be a permutation of range(arity)")
        self.iterset, self.toset, self.arity = map_.iterset, map_.toset, map_.arity
        self.name = f"permuted_{map_.name}"

    @property
    def offset(self):
        return None if self.map_.offset is None else np.asarray(self.map_.offset)[self.permutation]

    @property
    def offset_quotient(self):
        oq = getattr(self.map_, "offset_quotient", None)
        return None if oq is None else np.asarray(oq)[self.permutation]


def _base(map_):
    return map_.map_ if isinstance(map_, PermutedMap) else map_


def distinct_maps(args):
    """The parloop's map arguments: distinct by identity, first-use order; a
    PermutedMap contributes its base map (pyop2/parloop.py:203-212)."""
    maps = []
    for a in args:
        for m in (getattr(a, "map", None), getattr(a, "cmap", None)):
            if m is not None and not any(_base(m) is b for b in maps):
                maps.append(_base(m))
    return maps


class WrapperSpec:
    """The compile-time description of one generic parloop."""

    def __init__(self, kernel: CStringKernel, args, *, extruded=False, subset=False,
                 iteration_region="ALL", interior_horizontal=None, pass_layer_arg=False,
                 extruded_periodic=False, constant_layers=True):
        from . import op2
        if interior_horizontal is None:
            # every indirect argument of an interior-horizontal-facet loop packs the cells
            # below and above the facet (pyop2/codegen/builder.py:840-844)
            interior_horizontal = iteration_region == "ON_INTERIOR_FACETS"
        self.kernel = kernel
        self.maps = distinct_maps(args)
        if len(args) > _lib.WRAP_MAX_ARGS or len(self.maps) > _lib.WRAP_MAX_MAPS:
            raise ValueError("too many arguments / maps for the generic wrapper")
        self._keep = []
        arr = (_lib.WrapperArg * len(args))()
        key = [kernel.code, kernel.name, bool(extruded), bool(subset), iteration_region,
               bool(interior_horizontal), bool(pass_layer_arg), bool(extruded_periodic), bool(constant_layers)]
        for i, a in enumerate(args):
            w = arr[i]
            w.access = int(a.access)
            w.map = w.map2 = -1
            w.interior_horizontal = int(bool(interior_horizontal) and getattr(a, "map", None) is not None)
            data = a.data
            if isinstance(data, op2.Mat):
                rmap, cmap = a.map, a.cmap
                w.kind, w.dtype = _lib.ARG_MAT, _lib.F64
                w.dim = w.dim2 = data.bs
                w.map, w.map2 = self._slot(rmap), self._slot(cmap)
                w.arity, w.arity2 = rmap.arity, cmap.arity
                w.offset, w.offset2 = self._ints(rmap.offset), self._ints(cmap.offset)
                roq, coq = getattr(rmap, "offset_quotient", None), getattr(cmap, "offset_quotient", None)
                w.offset_quotient, w.offset_quotient2 = self._ints(roq), self._ints(coq)
                key.append(("mat", w.access, data.bs, w.map, w.map2, w.arity, w.arity2,
                            self._tup(rmap.offset), self._tup(cmap.offset), self._tup(roq), self._tup(coq)))
            elif isinstance(data, op2.Global):
                w.kind = _lib.ARG_GLOBAL
                w.dtype = _DTYPE_CODE[np.dtype(data._data.dtype)]
                w.dim = int(np.prod(data.dim))
                key.append(("glob", w.access, w.dtype, w.dim))
            else:
                w.kind = _lib.ARG_DAT
                w.dtype = _DTYPE_CODE[np.dtype(data.dtype)]
                w.dim = data.cdim
                m = a.map
                if m is not None:
                    w.map, w.arity = self._slot(m), m.arity
                    w.offset = self._ints(m.offset)
                    oq = getattr(m, "offset_quotient", None)
                    w.offset_quotient = self._ints(oq)
                    perm = m.permutation if isinstance(m, PermutedMap) else None
                    w.permutation = self._ints(perm)
                    w.mixed_continuation = int(bool(getattr(a, "mixed_continuation", False)))
                    key.append(("dat", w.access, w.dtype, w.dim, w.map, w.arity, self._tup(m.offset),
                                self._tup(perm), self._tup(oq), w.mixed_continuation))
                else:
                    key.append(("dat", w.access, w.dtype, w.dim, -1))
        self.cache_key = tuple(key)
        d = _lib.WrapperDesc()
        d.kernel_source = kernel.code.encode()
        d.kernel_name = kernel.name.encode()
        d.nargs, d.args = len(args), arr
        d.extruded, d.subset = int(bool(extruded)), int(bool(subset))
        d.iteration_region = _REGIONS[iteration_region]
        d.pass_layer_arg = int(bool(pass_layer_arg))
        d.extruded_periodic = int(bool(extruded_periodic))
        d.variable_layers = int(bool(extruded) and not constant_layers)
        self._keep.append(arr)
        self.desc = d

    # -- helpers
    def _slot(self, m):
        for i, b in enumerate(self.maps):
            if b is _base(m):
                return i
        raise AssertionError("map not registered")

    def _ints(self, v):
        if v is None:
            return None
        a = np.ascontiguousarray(v, dtype=np.int32)
        self._keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int32))

    @staticmethod
    def _tup(v):
        return None if v is None else tuple(int(x) for x in np.asarray(v).ravel())

    # -- products (no GPU needed for the first two)
    def source(self) -> str:
        """The generated CUDA source of ``wrap_<name>``."""
        L = _lib.load()
        need = C.c_size_t()
        _lib.check(L.fdb_wrapper_source(C.byref(self.desc), None, 0, C.byref(need)), "fdb_wrapper_source")
        buf = C.create_string_buffer(need.value)
        _lib.check(L.fdb_wrapper_source(C.byref(self.desc), buf, need.value, C.byref(need)), "fdb_wrapper_source")
        return buf.value.decode()

    def compile(self) -> bytes:
        """NVRTC-compile for sm_100a; returns the cubin image (ahead-of-time path)."""
        L = _lib.load()
        need = C.c_size_t()
        _lib.check(L.fdb_wrapper_compile(C.byref(self.desc), None, 0, C.byref(need)), "fdb_wrapper_compile")
        buf = (C.c_char * need.value)()
        _lib.check(L.fdb_wrapper_compile(C.byref(self.desc), buf, need.value, C.byref(need)), "fdb_wrapper_compile")
        return bytes(buf)

    def create(self):
        """Generate, compile and load on the active device -> kernel handle."""
        h = C.c_void_p()
        _lib.check(_lib.lib().fdb_wrapper_create(C.byref(self.desc), C.byref(h)), "fdb_wrapper_create")
        return h


_handles = {}     # cache_key -> handle: the reference caches compiled global kernels the same way


def _handle(spec: WrapperSpec):
    h = _handles.get(spec.cache_key)
    if h is None:
        h = _handles[spec.cache_key] = spec.create()
    return h


def par_loop(kernel: CStringKernel, iterset, *args, iteration_region="ALL", location="device",
             interior_horizontal=None, pass_layer_arg=False):
    """``op2.par_loop(op2.Kernel(code, name), iterset, *args)`` for a C-string
    kernel, one generated wrapper per distinct argument description.
    ``location="device"``: Dats stay resident on the GPU; ``"host"``: the drop-in
    call with NumPy buffers (mirror cache keyed on ``dat_version``, every written
    Dat copied back).  Follows pyop2/parloop.py:243-260 without the halo phases
    (generic parloops run unpartitioned for now)."""
    from . import op2
    if kernel.accesses is not None and tuple(a.access for a in args) != tuple(kernel.accesses):
        raise ValueError("access descriptors do not match the kernel's")
    # a MixedDat argument is one local tensor for the kernel and one wrapper argument per block
    args = tuple(b for a in args for b in (a.split() if isinstance(a, op2.MixedArg) else (a,)))
    if kernel.accesses is not None and len(args) != len(kernel.accesses):
        kernel = CStringKernel(kernel.code, kernel.name, None, kernel.flop_count, kernel.opts)
    if location == "host":
        return _par_loop_host(kernel, iterset, args, iteration_region, interior_horizontal, pass_layer_arg)
    base = iterset.superset if isinstance(iterset, op2.Subset) else iterset
    for a in args:
        for m in (a.map, getattr(a, "cmap", None)):
            if m is None:
                continue
            if m.iterset is not base:
                raise op2.MapValueError(f"map {m.name} is not defined on the iteration set")
        if a.map is not None and not isinstance(a.data, op2.Mat) and a.map.toset is not a.data.dataset.set:
            raise op2.MapValueError(f"map {a.map.name} does not target {a.data.name}'s set")
        if a.map is None and isinstance(a.data, op2.Dat) and a.data.dataset.set is not base:
            raise op2.MapValueError(f"direct argument {a.data.name} is not defined on the iteration set")
    spec = WrapperSpec(kernel, args, extruded=base._extruded, subset=isinstance(iterset, op2.Subset),
                       iteration_region=iteration_region, interior_horizontal=interior_horizontal,
                       pass_layer_arg=pass_layer_arg,
                       extruded_periodic=getattr(base, "extruded_periodic", False),
                       constant_layers=getattr(base, "constant_layers", True))
    h = _handle(spec)
    ca = _lib.CallArgs()
    lgmat = []
    ptrs = []
    for a in args:
        if isinstance(a.data, op2.Mat):
            ptrs.append(a.data.handle.value)
            if a.lgmaps is not None:
                lgmat.append(a)
        elif isinstance(a.data, op2.Global):
            ptrs.append(a.data._data.ctypes.data)
        else:
            ptrs.append(a.data.device_ptr)
    subset = None
    if isinstance(iterset, op2.Subset):
        if not hasattr(iterset, "_dev_idx"):
            iterset._dev_idx = op2.DeviceArray.from_host(iterset.indices)
        subset = iterset._dev_idx.ptr
    layers = base.layers_array.ravel() if base._extruded else None
    L = _lib.lib()
    for a in lgmat:
        r, c = (np.ascontiguousarray(v, dtype=np.int32) for v in a.lgmaps)
        _lib.check(L.fdb_mat_set_lgmaps(a.data.handle, r.ctypes.data, c.ctypes.data))
    # distributed protocol (pyop2/parloop.py:243-260, 354-455): ghost refresh of read Dats
    # overlapped with the core part, ghost contributions of INC Dats summed into their
    # owners afterwards, Globals reduced over the ranks
    nranks = L.fdb_comm_size()
    reads = [a.data for a in args if a.access in (op2.READ, op2.RW) and isinstance(a.data, op2.Dat)
             and a.data.dataset.halo is not None and not a.data.halo_valid and a.map is not None]
    incs = [a.data for a in args if a.access == op2.INC and isinstance(a.data, op2.Dat)
            and a.data.dataset.halo is not None and not a.data.frozen_halo]
    gouts = [a for a in args if isinstance(a.data, op2.Global) and a.access != op2.READ]
    saved = {}
    if nranks > 1:
        for a in gouts:
            if a.data._data.dtype != np.float64:
                raise NotImplementedError("distributed reductions of non-float64 Globals")
            if a.access == op2.INC:          # privatise: local sum from zero, added after the all-reduce
                saved[id(a)] = a.data._data.copy()
                a.data._data[...] = 0
    # ghost rows of accumulated Dats start from the identity of the reduction
    # (pyop2/types/dat.py:633-636), so that only this loop's contributions travel back
    for a in args:
        if (a.access in (op2.INC, op2.MIN, op2.MAX) and isinstance(a.data, op2.Dat)
                and a.data.dataset.halo is not None and not a.data.frozen_halo):
            a.data._reset_ghost_rows(a.access)
    for d in reads:
        d.dataset.halo.global_to_local_begin(d)
    first = True
    try:
        for start, end in (iterset.core_part, iterset.owned_part):
            if not first:
                for d in reads:
                    d.dataset.halo.global_to_local_end(d)
                reads = []
            first = False
            if end <= start:
                continue
            ca.start, ca.end = int(start), int(end)
            if layers is not None:
                ca.layers = layers.ctypes.data_as(C.POINTER(C.c_int32))
                if not base.constant_layers:
                    ca.layers_count, ca.layers_version = len(base.layers_array), base._generation
            ca.subset = subset
            ca.nargs, ca.args = len(ptrs), (C.c_void_p * len(ptrs))(*ptrs)
            mp = [m.device_ptr for m in spec.maps]
            ca.nmaps, ca.maps = len(mp), (C.c_void_p * max(len(mp), 1))(*mp)
            ca.location, ca.writeback, ca.output_is_zero = _lib.LOC_DEVICE, 0, 0
            _lib.check(L.fdb_kernel_call(h, C.byref(ca)), "wrap_" + kernel.name)
    finally:
        for a in lgmat:
            _lib.check(L.fdb_mat_set_lgmaps(a.data.handle, None, None))
    for a in args:
        if a.access == op2.READ:
            continue
        if isinstance(a.data, op2.Dat):
            a.data._device_written()
            a.data.halo_valid = False
        else:
            a.data.dat_version += 1
    for d in incs:
        d.dataset.halo.local_to_global_begin(d)
        d.dataset.halo.local_to_global_end(d)
    if nranks > 1:
        for a in gouts:
            buf = op2.DeviceArray.from_host(a.data._data)
            op = {op2.INC: 0, op2.MIN: 1, op2.MAX: 2}[a.access]
            _lib.check(L.fdb_allreduce(buf.ptr, a.data._data.size, op), "fdb_allreduce")
            buf.to_host(a.data._data)
            if a.access == op2.INC:
                a.data._data[...] += saved[id(a)]
    return spec


def _par_loop_host(kernel, iterset, args, iteration_region, interior_horizontal=None, pass_layer_arg=False):
    """The arglist of pyop2/parloop.py:203-212 with host pointers, sizes and versions."""
    from . import op2
    base = iterset.superset if isinstance(iterset, op2.Subset) else iterset
    if any(isinstance(a.data, op2.Mat) for a in args):
        raise NotImplementedError("host-pointer mode takes Dats and Globals (Mats live on the device)")
    spec = WrapperSpec(kernel, args, extruded=base._extruded, subset=isinstance(iterset, op2.Subset),
                       iteration_region=iteration_region, interior_horizontal=interior_horizontal,
                       pass_layer_arg=pass_layer_arg,
                       extruded_periodic=getattr(base, "extruded_periodic", False),
                       constant_layers=getattr(base, "constant_layers", True))
    h = _handle(spec)
    for a in args:
        if isinstance(a.data, op2.Dat):
            a.data._sync_host()
    ptrs = [a.data._data.ctypes.data for a in args]
    nbytes = [a.data._data.nbytes for a in args]
    vers = [a.data.dat_version for a in args]
    maps = [m.values_with_halo for m in spec.maps]
    layers = base.layers_array.ravel() if base._extruded else None
    ca = _lib.CallArgs()
    for start, end in (iterset.core_part, iterset.owned_part):
        if end <= start:
            continue
        ca.start, ca.end = int(start), int(end)
        if layers is not None:
            ca.layers = layers.ctypes.data_as(C.POINTER(C.c_int32))
            if not base.constant_layers:
                ca.layers_count, ca.layers_version = len(base.layers_array), base._generation
        ca.subset = iterset.indices.ctypes.data if isinstance(iterset, op2.Subset) else None
        ca.nargs, ca.args = len(ptrs), (C.c_void_p * len(ptrs))(*ptrs)
        ca.arg_bytes = (C.c_size_t * len(ptrs))(*nbytes)
        ca.arg_versions = (C.c_uint64 * len(ptrs))(*vers)
        ca.nmaps = len(maps)
        ca.maps = (C.c_void_p * max(len(maps), 1))(*[m.ctypes.data for m in maps])
        ca.map_bytes = (C.c_size_t * max(len(maps), 1))(*[m.nbytes for m in maps])
        ca.map_versions = (C.c_uint64 * max(len(maps), 1))(*[getattr(m, "_generation", 0) for m in spec.maps])
        ca.subset_version = iterset._generation if isinstance(iterset, op2.Subset) else 0
        ca.location, ca.writeback, ca.output_is_zero = _lib.LOC_HOST, 1, 0
        _lib.check(_lib.lib().fdb_kernel_call(h, C.byref(ca)), "wrap_" + kernel.name)
        # the engine recorded version+1 for every written mirror (pyop2/parloop.py:262-272)
        for i, a in enumerate(args):
            if a.access != op2.READ and isinstance(a.data, op2.Dat):
                vers[i] += 1
    for a in args:
        if a.access == op2.READ:
            continue
        if isinstance(a.data, op2.Dat):
            a.data.increment_dat_version()
            a.data._host_valid, a.data._dev_valid, a.data._is_zero = True, False, False
            a.data.halo_valid = False
        else:
            a.data.dat_version += 1
    return spec
