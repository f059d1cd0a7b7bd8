"""Host-side behaviour of the PyOP2 mirror that needs no GPU: set partitions,
map validation, Dat versioning / lazy zero, parloop argument checks -- the
Python-level errors the reference raises before any launch
(pyop2/parloop.py:175-189, 472-501)."""
import numpy as np
import pytest

from firedrake_b200 import op2


def test_set_partitions():
    s = op2.Set((3, 5, 8))
    assert s.core_part == (0, 3) and s.owned_part == (3, 5) and s.total_size == 8
    e = op2.ExtrudedSet(s, layers=11)
    assert e._extruded and e.layers == 11 and e.layers_array.tolist() == [[0, 11]]
    assert e.sizes == s.sizes
    with pytest.raises(ValueError):
        op2.Set((5, 3, 8))
    with pytest.raises(ValueError):
        op2.ExtrudedSet(s, layers=1)
    sub = op2.Subset(s, [7, 1, 4, 1])
    assert sub.indices.tolist() == [1, 4, 7] and sub.sizes == (1, 2, 3)
    with pytest.raises(ValueError):
        op2.Subset(s, [9])


def test_map_validation():
    cells, nodes = op2.Set(2), op2.Set(4)
    m = op2.Map(cells, nodes, 3, [0, 1, 3, 2, 3, 1])
    assert m.values.shape == (2, 3)
    with pytest.raises(op2.MapValueError):
        op2.Map(cells, nodes, 3, [0, 1, 3, 2, 3, 4])          # out of range
    with pytest.raises(op2.MapValueError):
        op2.Map(cells, nodes, 3, [0, 1, 3])                    # wrong number of rows
    with pytest.raises(op2.MapValueError):
        op2.Map(cells, nodes, 3, [0, 1, 3, 2, 3, 1], offset=[1, 1])


def test_dat_versioning_and_lazy_zero():
    nodes = op2.Set(5)
    d = op2.Dat(nodes, np.arange(5.0))
    v0 = d.dat_version
    assert d.data_ro.tolist() == [0, 1, 2, 3, 4] and d.dat_version == v0     # read-only: no bump
    d.data[2] = 7.0
    assert d.dat_version == v0 + 1 and not d.halo_valid
    with pytest.raises(ValueError):
        d.data_ro[0] = 1.0
    d.zero()
    assert d._is_zero and not d._host_valid                                   # nothing touched yet
    assert d.data_ro.tolist() == [0, 0, 0, 0, 0]                              # materialises on read
    vec = op2.Dat(op2.DataSet(nodes, 3))
    assert vec.data_ro.shape == (5, 3) and vec.cdim == 3
    d2 = op2.Dat(op2.Set((3, 3, 5)), np.arange(5.0))
    assert d2.data_ro.shape == (3,) and d2.data_ro_with_halos.shape == (5,)   # ghosts at the tail


def test_kernel_descriptors():
    k = op2.Kernel("helmholtz", degree=3)
    assert k.accesses == (op2.INC, op2.READ, op2.READ) and k.name == "form0_cell_integral"
    k2 = op2.Kernel("helmholtz", degree=2, rank=2)
    assert k2.accesses == (op2.INC, op2.READ) and k2.name == "form00_cell_integral"
    kd = op2.Kernel("helmholtz", degree=2, diagonal=True)
    assert kd.accesses == (op2.INC, op2.READ)
    kf = op2.Kernel("dg_advection", integral="interior_facet")
    assert len(kf.accesses) == 6 and kf.name == "form0_interior_facet_integral"
    assert op2.GlobalKernel(k, []).name == "wrap_form0_cell_integral"        # global_kernel.py:344-346


def test_parloop_argument_checks():
    cells = op2.ExtrudedSet(op2.Set(2), 3)
    nodes, verts = op2.Set(12), op2.Set(12)
    m0 = op2.Map(cells, nodes, 8, np.arange(16) % 12, offset=np.ones(8, dtype=np.int32))
    m1 = op2.Map(cells, verts, 8, np.arange(16) % 12, offset=np.ones(8, dtype=np.int32))
    x, y = op2.Dat(nodes), op2.Dat(nodes)
    X = op2.Dat(op2.DataSet(verts, 3))
    k = op2.Kernel("helmholtz", degree=1)
    gk = op2.GlobalKernel(k, [m0, m1], extruded=True)
    op2.Parloop(gk, cells, [y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0)])          # fine
    with pytest.raises(ValueError):                                                     # wrong access
        op2.Parloop(gk, cells, [y(op2.READ, m0), X(op2.READ, m1), x(op2.READ, m0)])
    with pytest.raises(ValueError):                                                     # arity
        op2.Parloop(gk, cells, [y(op2.INC, m0), X(op2.READ, m1)])
    with pytest.raises(op2.MapValueError):                                              # map/dat mismatch
        op2.Parloop(gk, cells, [y(op2.INC, m1), X(op2.READ, m1), x(op2.READ, m0)])
    other = op2.ExtrudedSet(op2.Set(2), 3)
    with pytest.raises(op2.MapValueError):                                              # map on another set
        op2.Parloop(gk, other, [y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0)])
    with pytest.raises(NotImplementedError):
        op2.GlobalKernel(k, [m0, m1], extruded=True, constant_layers=False)


def test_sparsity_restrictions():
    cells, nodes = op2.Set(2), op2.Set(4)
    m = op2.Map(cells, nodes, 3, [0, 1, 3, 2, 3, 1])
    s = op2.Sparsity((nodes, nodes), [(m, m, None)])
    assert s.shape == (4, 4)
    sb = op2.Sparsity((op2.DataSet(nodes, 2), op2.DataSet(nodes, 2)), [(m, m, None)])
    assert sb.bs == 2 and sb.shape == (4, 4)                    # node-level pattern, 2x2 blocks
    with pytest.raises(NotImplementedError):                    # rectangular blocks
        op2.Sparsity((op2.DataSet(nodes, 2), op2.DataSet(nodes, 3)), [(m, m, None)])
    m2 = op2.Map(cells, nodes, 3, [0, 1, 3, 2, 3, 1])
    with pytest.raises(NotImplementedError):
        op2.Sparsity((nodes, nodes), [(m, m2, None)])


def test_quadmesh_slabs():
    from firedrake_b200.assemble import dg_slab
    from firedrake_b200.utility_meshes import QuadMesh
    g = QuadMesh(9, 4)
    assert g.nbr.shape == (36, 4) and (g.nbr[0] == [-1, 4, -1, 1]).all()
    owned = 0
    for r in range(3):
        m, neigh = dg_slab(9, 4, r, 3)
        owned += m.num_owned_cells
        for _, send, recv in neigh:
            assert len(send) == len(recv) == 16
            assert send.max() < 4 * m.num_owned_cells <= recv.min()      # send owned, receive into the tail
        # owned cells see a neighbour everywhere except on the global boundary
        gc = m.cell_global_column[:m.num_owned_cells]
        assert ((m.nbr[:m.num_owned_cells, 0] < 0) == (gc == 0)).all()
        assert ((m.nbr[:m.num_owned_cells, 1] < 0) == (gc == 8)).all()
    assert owned == 36


def test_firedrake_hook_needs_firedrake():
    """The real-Firedrake hook imports cleanly and fails loudly (ImportError)
    where PyOP2 is absent -- it can never silently replace anything here."""
    import firedrake_b200.firedrake_hook as hook
    with pytest.raises(ImportError):
        hook.install()
    hook.uninstall()        # no-op when nothing was installed


def test_composed_map_is_materialised():
    """op2.ComposedMap (pyop2/types/map.py:219-279): facets -> cells -> nodes; the inner maps have arity 1."""
    rng = np.random.default_rng(3)
    facets, cells, nodes = op2.Set(11), op2.Set(7), op2.Set(20)
    c2n = op2.Map(cells, nodes, 4, rng.integers(0, 20, (7, 4)), offset=None)
    f2c = op2.Map(facets, cells, 1, rng.integers(0, 7, (11, 1)))
    cm = op2.ComposedMap(c2n, f2c)
    assert cm.iterset is facets and cm.toset is nodes and cm.arity == 4 and cm.maps_ == (c2n, f2c)
    assert np.array_equal(cm.values, c2n.values[f2c.values[:, 0]])
    sub = op2.Set(5)
    s2f = op2.Map(sub, facets, 1, rng.integers(0, 11, (5, 1)))
    cm3 = op2.ComposedMap(c2n, f2c, s2f)
    assert np.array_equal(cm3.values, c2n.values[f2c.values[s2f.values[:, 0], 0]])
    with pytest.raises(op2.MapValueError):
        op2.ComposedMap(f2c, c2n)                      # c2n lands on nodes, f2c iterates facets
    with pytest.raises(op2.MapValueError):
        op2.ComposedMap(c2n, op2.Map(facets, cells, 2, rng.integers(0, 7, (11, 2))))   # inner arity 2
    # extruded: the offset of the outer map is inherited
    ecells = op2.ExtrudedSet(op2.Set(7), 4)
    ec2n = op2.Map(ecells, nodes, 2, rng.integers(0, 10, (7, 2)), offset=[1, 1])
    efac = op2.ExtrudedSet(op2.Set(11), 4)
    ef2c = op2.Map(efac, ecells, 1, rng.integers(0, 7, (11, 1)), offset=[0])
    ecm = op2.ComposedMap(ec2n, ef2c)
    assert np.array_equal(ecm.offset, [1, 1]) and ecm.iterset is efac


def test_subset_nesting_and_set_algebra():
    """Subsets of subsets address the parent (pyop2/types/set.py:413-416); intersection / union /
    difference / symmetric_difference work on the index lists (set.py:486-547)."""
    s = op2.Set((6, 10, 14))
    a = op2.Subset(s, [1, 3, 5, 7, 9, 11, 13])
    b = a([0, 2, 2, 6])                                   # positions in a -> entries 1, 5, 13 of s
    assert b.superset is s and list(b.indices) == [1, 5, 13] and b.sizes == (2, 2, 3)
    assert list(a.owned_indices) == [1, 3, 5, 7, 9]
    c = op2.Subset(s, [0, 1, 2, 3, 13])
    assert list(a.intersection(c).indices) == [1, 3, 13] and a.intersection(s) is a
    assert list(a.union(c).indices) == [0, 1, 2, 3, 5, 7, 9, 11, 13] and a.union(s) is s
    assert list(a.difference(c).indices) == [5, 7, 9, 11] and len(a.difference(s).indices) == 0
    assert list(a.symmetric_difference(c).indices) == [0, 2, 5, 7, 9, 11]
    assert list(a.symmetric_difference(s).indices) == [0, 2, 4, 6, 8, 10, 12]
    with pytest.raises(TypeError):
        a.union(op2.Subset(op2.Set(14), [1]))
    with pytest.raises(ValueError):
        a([7])


def test_global_vector_operations():
    """The host-side vector interface of op2.Global (pyop2/types/glob.py:33-180)."""
    g = op2.Global(3, [1.0, 2.0, 3.0])
    h = g.duplicate()
    assert h is not g and np.array_equal(h.data_ro, g.data_ro) and g.shape == (3,) and g.nbytes == 24
    v = g.dat_version
    g.axpy(2.0, h)
    assert g.dat_version > v and np.array_equal(g.data_ro, [3.0, 6.0, 9.0])
    g.maxpy([1.0, -1.0], [h, h])
    assert np.array_equal(g.data_ro, [3.0, 6.0, 9.0]) and g.inner(h) == 3 + 12 + 27
    g.copy(h)
    assert np.array_equal(h.data_ro, g.data_ro)
    g.zero()
    assert not g.data_ro.any() and g.split() == (g,)
