"""Slab partition of the synthetic extruded mesh across ranks (host side).

The reference partitions the BASE mesh with DMPlex and every layer of a column
stays on one rank (firedrake/mesh.py:1138-1179, 1813-1820); entities are
labelled core / owned / ghost and sets are ordered ``[core | owned | ghost]``
(pyop2/types/set.py:38-52).  This module produces the same structure for a
1-D split along x:

* rank r holds base cells ``ix in [x0, x1)``; it iterates its OWNED cells only
  (INC loops do not run on halo cells; contributions to ghost dofs go back to
  the owner through the local->global reduce, pyop2/parloop.py:255-260);
* the dof columns on the slab's LEFT face belong to rank r-1 (ghost tail of
  rank r); the columns on the RIGHT face are owned by r and are ghosts on r+1;
* cells touching ghost dofs come last in the cell order: ``core_part`` can be
  computed while the global->local exchange is in flight.
"""
from __future__ import annotations

import numpy as np

from .utility_meshes import ExtrudedHexMesh


def slab_bounds(nx, nranks, rank):
    base, rem = divmod(nx, nranks)
    x0 = rank * base + min(rank, rem)
    return x0, x0 + base + (1 if rank < rem else 0)


class SlabPartition:
    def __init__(self, nx, ny, nz, degree, rank, nranks, warp=0.0, Lx=1.0, Ly=1.0, Lz=1.0):
        if nranks > nx:
            raise ValueError("more ranks than base-cell columns")
        self.rank, self.nranks = rank, nranks
        x0, x1 = slab_bounds(nx, nranks, rank)
        self.x0, self.x1 = x0, x1
        self.mesh = ExtrudedHexMesh(x1 - x0, ny, nz, Lx=Lx, Ly=Ly, Lz=Lz, warp=warp, ix0=x0,
                                    nx_global=nx, ghost_left=rank > 0)
        self.V = self.mesh.function_space(degree)
        self.neighbours = self.halo_lists(self.V)
        self.coord_neighbours = self.halo_lists(self.mesh.coord_space)

    def halo_lists(self, V):
        """(rank, send, recv) per neighbour for function space V on this slab."""
        out = []
        empty = np.zeros(0, dtype=np.int32)
        if self.rank > 0:                      # my left face is owned by rank-1
            out.append((self.rank - 1, empty, V.plane_nodes(0)))
        if self.rank < self.nranks - 1:        # my right face is a ghost on rank+1
            out.append((self.rank + 1, V.plane_nodes(self.mesh.nx), empty))
        return out

    @property
    def cell_sizes(self):
        """(core, owned, total) sizes of the column set."""
        return (self.mesh.num_core_cells, self.mesh.num_base_cells, self.mesh.num_base_cells)

    @property
    def node_sizes(self):
        V = self.V
        return (V.owned_node_count, V.owned_node_count, V.node_count)
