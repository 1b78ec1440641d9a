"""One tiny matvec of the hot path on cuda:0, checked against the oracle (used by
__graft_entry__.smoke() and tests/test_gpu_parity.py)."""
import numpy as np


def run(dkt):
    import cases
    import flat
    dim, md = 4, 10
    xyz, lev = dkt.trees.moving_ball_tree(dim, 5, md)
    da = dkt.DA(xyz, lev, dim, 1, md)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    nx, nl = da.nodes()
    assert np.array_equal(nx, t.node_xyz) and np.array_equal(nl, t.node_lev), "node order differs from the oracle"
    K = flat.laplace_kref(dim, 1)
    u = cases.input_vector(da.n_nodes)
    v = da.matvec(dkt.Operator.dense(K, alpha=dim - 2.0), u)
    vo = flat.matvec(t, u, K, alpha=dim - 2.0)
    err = np.abs(v - vo).max() / np.abs(vo).max()
    assert err < 1e-12, err
    print("smoke ok: 4-D p=1 moving ball, %d elements, %d nodes, class %s, rel err %.2e, %d kernel launches"
          % (da.n_elem, da.n_nodes, da.tree_class, err, dkt.kernel_launch_count()))
    da.close()
