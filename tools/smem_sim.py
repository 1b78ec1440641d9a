#!/usr/bin/env python
"""Shared-memory bank-conflict model of the chunked matvec tables (analysis tool, CPU only).

Rebuilds the REGULAR chunk set the way dkt_chunks.cu does (SFC chunks of 256 elements, XOR slot
schedule, jagged-diagonal positions padded to k mod 16) from the oracle's element->node table and
counts the shared-memory wavefronts of the two random accesses of phase T2:
    gather   LDS.64 un[n]                 (bank pair = n mod 16, equal addresses broadcast)
    scatter  STS.64 X[jd[k] + n]          (bank pair = position mod 16)
A 64-bit warp access is served as two half-warp transactions; each costs max-over-bank-pairs of the
number of DISTINCT 8-byte words, so the ideal is 2 wavefronts per instruction.  Use it to evaluate
alternative assignments of node ranks / run positions before touching the CUDA build kernel.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dendro-kt_b200")):
    sys.path.insert(0, p)
import flat  # noqa: E402
import dkt  # noqa: E402

E = 256


def regular_chunks(t, max_depth):
    """Yield (e2n_chunk[ne,16] permuted by the XOR schedule) for the regular elements, in tree order."""
    reg = np.nonzero((t.e2n >= 0).all(axis=1))[0]
    e2n = t.e2n[reg]
    lev = t.mv_lev[reg].astype(np.int64)
    xyz = t.mv_xyz[reg].astype(np.int64)
    c = np.zeros(len(reg), dtype=np.int64)
    for d in range(t.dim):
        c |= ((xyz[:, d] >> (max_depth - lev)) & 1) << d
    s = np.arange(t.N)[None, :]
    perm = s ^ c[:, None]                      # slot s holds rank s ^ c
    e2n_perm = np.take_along_axis(e2n, perm, axis=1)
    for a in range(0, len(reg), E):
        yield e2n_perm[a:a + E]


def default_assignment(slots):
    """dkt_chunks.cu: nodes ranked by (run length desc, id asc); k = order of the slot inside its run in
    element-major slot order; diagonal k starts at a position == k (mod 16) for k < 16."""
    ne, N = slots.shape
    flat_nodes = slots.ravel()                 # slot index = e*N + s (element-major)
    order = np.argsort(flat_nodes, kind="stable")
    sorted_nodes = flat_nodes[order]
    uniq, start, cnt = np.unique(sorted_nodes, return_index=True, return_counts=True)
    node_of_sorted = np.repeat(np.arange(len(uniq)), cnt)
    k_sorted = np.arange(len(sorted_nodes)) - start[node_of_sorted]
    rank_order = np.lexsort((uniq, -cnt))      # len desc, id asc
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[rank_order] = np.arange(len(uniq))
    loc = np.empty(ne * N, dtype=np.int64)
    k = np.empty(ne * N, dtype=np.int64)
    loc[order] = rank[node_of_sorted]
    k[order] = k_sorted
    return loc.reshape(ne, N), k.reshape(ne, N), cnt[rank_order]


def positions(loc, k, lens_by_rank):
    maxlen = int(lens_by_rank.max())
    count_k = np.array([(lens_by_rank > j).sum() for j in range(maxlen)])
    jd = np.zeros(maxlen + 1, dtype=np.int64)
    acc = 0
    for j in range(maxlen):
        if j < 16:
            while (acc & 15) != j:
                acc += 1
        jd[j] = acc
        acc += count_k[j]
    return jd[k] + loc


def wavefronts(addr):
    """addr: (ne, N) 8-byte word addresses, one instruction per column per warp of 32 elements."""
    ne, N = addr.shape
    total, instr = 0, 0
    for w0 in range(0, ne, 32):
        blk = addr[w0:w0 + 32]
        for s in range(N):
            for h0 in (0, 16):
                a = np.unique(blk[h0:h0 + 16, s])
                if len(a) == 0:
                    continue
                total += np.bincount(a % 16, minlength=16).max()
            instr += 1
    return total, instr


def main():
    dim, md = 4, 10
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    xyz, lev = dkt.trees.moving_ball_tree(dim, level, md)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    g = s = ni = 0
    nodes = elems = 0
    for slots in regular_chunks(t, md):
        loc, k, lens = default_assignment(slots)
        pos = positions(loc, k, lens)
        wg, n1 = wavefronts(loc)
        ws, _ = wavefronts(pos)
        g, s, ni = g + wg, s + ws, ni + n1
        nodes += len(lens)
        elems += len(slots)
    print("regular elements %d, chunk nodes/element %.2f" % (elems, nodes / elems))
    print("gather  LDS.64: %.2f wavefronts/instruction (ideal 2)" % (g / ni))
    print("scatter STS.64: %.2f wavefronts/instruction (ideal 2)" % (s / ni))


if __name__ == "__main__":
    main()
