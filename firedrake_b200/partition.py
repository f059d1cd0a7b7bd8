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

``exec_halo=True`` (SURVEY.md section 8e option (ii)): every rank but the last additionally
holds an EXEC-HALO copy of its right neighbour's first cell column and executes it redundantly;
all its owned dofs then receive complete sums locally and the local->global reduce disappears.
The price is a wider ghost-read region (the dofs of that column, refreshed from the right
neighbour in the same exchange as the left ghost plane) and 1/width more element work.
"""
from __future__ import annotations

import numpy as np

from .utility_meshes import ExtrudedHexMesh


def slab_bounds(nx, nranks, rank):
    base, rem = divmod(nx, nranks)
    x0 = rank * base + min(rank, rem)
    return x0, x0 + base + (1 if rank < rem else 0)


class SlabPartition:
    def __init__(self, nx, ny, nz, degree, rank, nranks, warp=0.0, Lx=1.0, Ly=1.0, Lz=1.0, exec_halo=False):
        if nranks > nx:
            raise ValueError("more ranks than base-cell columns")
        self.rank, self.nranks = rank, nranks
        x0, x1 = slab_bounds(nx, nranks, rank)
        self.x0, self.x1 = x0, x1
        self.exec_halo = bool(exec_halo) and nranks > 1
        self.halo_right = self.exec_halo and rank < nranks - 1
        self.mesh = ExtrudedHexMesh(x1 - x0 + int(self.halo_right), ny, nz, Lx=Lx, Ly=Ly, Lz=Lz, warp=warp,
                                    ix0=x0, nx_global=nx, ghost_left=rank > 0, halo_right=self.halo_right)
        self.V = self.mesh.function_space(degree)
        self.neighbours = self.halo_lists(self.V)
        self.coord_neighbours = self.halo_lists(self.mesh.coord_space)

    def halo_lists(self, V):
        """(rank, send, recv) per neighbour for function space V on this slab."""
        out = []
        empty = np.zeros(0, dtype=np.int32)
        w = self.x1 - self.x0                  # owned cell columns; local plane w = my right face
        if self.rank > 0:                      # my left face is owned by rank-1 ...
            # ... which, in exec-halo mode, also reads the dofs of my first column
            send = V.column_region_nodes(0) if self.exec_halo else empty
            out.append((self.rank - 1, send, V.plane_nodes(0)))
        if self.rank < self.nranks - 1:        # my right face is a ghost on rank+1
            recv = V.column_region_nodes(w) if self.halo_right else empty
            out.append((self.rank + 1, V.plane_nodes(w), recv))
        return out

    @property
    def cell_sizes(self):
        """(core, owned, total) sizes of the column set."""
        return (self.mesh.num_core_cells, self.mesh.num_owned_cells, self.mesh.num_base_cells)

    @property
    def node_sizes(self):
        V = self.V
        return (V.owned_node_count, V.owned_node_count, V.node_count)
