"""Additive Schwarz patch smoother on the device.

Mirrors the pieces of the reference a patch preconditioner is made of:

* ``vertex_star_patches`` -- the dof sets ``firedrake.ASMStarPC`` builds with
  ``construct_dim = 0`` (firedrake/preconditioners/asm.py:150-230): for every mesh vertex, the dofs
  of all entities in its OPEN star (the vertex, and every edge / face / cell that contains it);
* ``PatchASM`` -- TinyASM's ``BlockJacobi`` (tinyasm/tinyasm.cpp:27-120): ``update`` extracts the
  dense patch blocks of the assembled operator and inverts them, ``apply`` adds
  ``inv(A[d_p, d_p]) b[d_p]`` into ``x[d_p]`` for every patch.  Both run in ``csrc/patch_asm.cu``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, op2


def vertex_star_patches(V, exclude=()):
    """(patch_ptr, patch_dofs) of the vertex-star patches of the scalar space ``V.V`` (an
    ``ExtrudedFunctionSpace``): on the p-refined lattice a dof belongs to the open star of the
    vertex at lattice position ``p * (i, j, k)`` iff it is closer than ``p`` to it in every
    direction.  ``exclude``: node indices to leave out of every patch (Dirichlet rows).
    Small / medium meshes (uses ``dof_lattice``)."""
    fs = V.V
    p = fs.degree
    lat = fs.dof_lattice()
    lo, hi = lat.min(axis=0), lat.max(axis=0)
    skip = np.zeros(fs.node_count, dtype=bool)
    skip[np.asarray(exclude, dtype=np.int64)] = True
    ptr, dofs = [0], []
    for i in range(lo[0], hi[0] + 1, p):
        in_i = np.abs(lat[:, 0] - i) < p
        for j in range(lo[1], hi[1] + 1, p):
            in_ij = in_i & (np.abs(lat[:, 1] - j) < p)
            for k in range(lo[2], hi[2] + 1, p):
                sel = np.nonzero(in_ij & (np.abs(lat[:, 2] - k) < p) & ~skip)[0]
                if len(sel):
                    dofs.append(sel)
                    ptr.append(ptr[-1] + len(sel))
    return (np.asarray(ptr, dtype=np.int64),
            np.concatenate(dofs).astype(np.int32) if dofs else np.zeros(0, dtype=np.int32))


class PatchASM:
    """``PatchASM(mat, patch_ptr, patch_dofs)``: additive Schwarz over dof patches of the assembled
    matrix ``mat`` (an ``op2.Mat``).  ``update()`` after (re)assembly, ``apply(b, x)`` computes
    ``x = sum_p R_p^T inv(R_p A R_p^T) R_p b`` on device-resident Dats."""

    def __init__(self, mat: op2.Mat, patch_ptr, patch_dofs):
        self.mat = mat
        self.ptr = np.ascontiguousarray(patch_ptr, dtype=np.int64)
        self.dofs = np.ascontiguousarray(patch_dofs, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(_lib.lib().fdb_asm_create(len(self.ptr) - 1, self.ptr.ctypes.data, self.dofs.ctypes.data,
                                             C.byref(h)), "fdb_asm_create")
        self._handle = h
        self.update()

    def update(self):
        ns = C.c_int()
        _lib.check(_lib.lib().fdb_asm_update(self._handle, self.mat.handle, C.byref(ns)), "fdb_asm_update")
        if ns.value:
            raise np.linalg.LinAlgError(f"{ns.value} singular patch block(s)")

    def apply(self, b: op2.Dat, x: op2.Dat):
        x.zero()
        _lib.check(_lib.lib().fdb_asm_apply(self._handle, b.device_ptr, x.device_ptr), "fdb_asm_apply")
        x._device_written()
        return x

    def inverse_blocks(self):
        """The inverted patch blocks, one (n_p, n_p) array per patch (tests)."""
        n = np.diff(self.ptr)
        out = np.empty(int((n * n).sum()))
        _lib.check(_lib.lib().fdb_asm_get_blocks(self._handle, out.ctypes.data), "fdb_asm_get_blocks")
        off = np.concatenate([[0], np.cumsum(n * n)])
        return [out[off[k]:off[k + 1]].reshape(n[k], n[k]) for k in range(len(n))]

    def __del__(self):
        try:
            if self._handle is not None and _lib._initialised is not None:
                _lib._lib.fdb_asm_destroy(self._handle)
        except Exception:
            pass
