"""NumPy model of the multi-GPU partition used by dkt_dist.cu (test infrastructure): SFC-contiguous
element ranges, node ownership by the first touching element, ghost lists.  Lets the N>1 exchange
logic run on CPU (gloo) against the oracle."""
import numpy as np


def partition(t, nranks):
    """t: oracle FlatTables.  Returns per-rank dicts with the element subset, the local node list
    [owned | ghosts grouped by owner], and send/recv lists (global ids, ascending)."""
    n_mv = len(t.mv_lev)
    bounds = [(n_mv * p) // nranks for p in range(nranks + 1)]
    rank_of_elem = np.searchsorted(np.array(bounds[1:]), np.arange(n_mv), side="right")
    n_nodes = len(t.node_lev)
    minsrc = np.full(n_nodes, n_mv, dtype=np.int64)
    refmask = np.zeros((n_nodes, nranks), dtype=bool)
    hang_row = {int(e): i for i, e in enumerate(t.hang_idx)}
    for e in range(n_mv):
        ids = t.e2n[e][t.e2n[e] >= 0]
        if e in hang_row:
            p = t.pnode[hang_row[e]]
            ids = np.concatenate([ids, p[p >= 0]])
        np.minimum.at(minsrc, ids, e)
        refmask[ids, rank_of_elem[e]] = True
    owner = rank_of_elem[minsrc]
    out = []
    for r in range(nranks):
        owned = np.nonzero(owner == r)[0]
        ghosts = [np.nonzero((owner == p) & refmask[:, r])[0] if p != r else np.zeros(0, dtype=np.int64) for p in range(nranks)]
        sends = [np.nonzero((owner == r) & refmask[:, p])[0] if p != r else np.zeros(0, dtype=np.int64) for p in range(nranks)]
        local = np.concatenate([owned] + ghosts)
        g2l = np.full(n_nodes, -1, dtype=np.int64)
        g2l[local] = np.arange(len(local))
        out.append(dict(elems=np.nonzero(rank_of_elem == r)[0], owned=owned, ghosts=ghosts, sends=sends, local=local, g2l=g2l))
    return out
