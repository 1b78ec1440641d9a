#!/usr/bin/env python
"""Prototype (CPU, NumPy) of a bank-aware assignment for the chunk tables, evaluated with the wavefront
model of tools/smem_sim.py.  Freedoms used (neither changes the matvec kernel):
  (F1) the rank n of a node inside its run-length class (jagged-diagonal storage only needs ranks sorted
       by run length) -> chooses the gather bank n mod 16 and shifts the scatter banks
  (F2) the order k of the writers inside a node's run -> scatter bank (k + n) mod 16 for k < 16
"""
import sys

import numpy as np

import smem_sim as sim


def groups_of(slots):
    """half-warp instruction groups: group id of every slot (ne, N) = ((e // 16) * N + s)."""
    ne, N = slots.shape
    e = np.arange(ne)[:, None]
    s = np.arange(N)[None, :]
    return (e // 16) * N + s


def bank_aware(slots):
    ne, N = slots.shape
    flat_nodes = slots.ravel()
    gid = groups_of(slots).ravel()
    ngroups = int(gid.max()) + 1
    uniq, inv, cnt = np.unique(flat_nodes, return_inverse=True, return_counts=True)
    nn = len(uniq)
    # run-length classes: ranks [a,b) per class, sorted by len desc
    order = np.lexsort((uniq, -cnt))
    lens_sorted = cnt[order]
    # capacity of every residue inside every class
    class_of_len = {}
    pos = 0
    free = {}
    for L in sorted(set(cnt.tolist()), reverse=True):
        m = int((cnt == L).sum())
        class_of_len[L] = (pos, pos + m)
        free[L] = {r: [x for x in range(pos, pos + m) if x % 16 == r] for r in range(16)}
        pos += m
    # slots of every node, groups of every node
    slot_lists = [[] for _ in range(nn)]
    for sidx, v in enumerate(inv):
        slot_lists[v].append(sidx)
    # ---- stage 1: residues for the gather ------------------------------------------------------------
    grp_res = np.zeros((ngroups, 16), dtype=np.int32)   # distinct nodes with that residue already in the group
    rank = np.full(nn, -1, dtype=np.int64)
    first_slot = np.array([sl[0] for sl in slot_lists])
    for v in np.argsort(first_slot, kind="stable"):
        L = int(cnt[v])
        gs = np.unique(gid[slot_lists[v]])
        cost = grp_res[gs].sum(axis=0).astype(np.float64)
        for r in range(16):
            if not free[L][r]:
                cost[r] = 1e9
        r = int(np.argmin(cost))
        rank[v] = free[L][r].pop(0)
        grp_res[gs, r] += 1
    # ---- stage 2: an independent column rank for X (scatter) + the writer order ------------------------
    free2 = {}
    for L, (a, b) in class_of_len.items():
        free2[L] = {r: [x for x in range(a, b) if x % 16 == r] for r in range(16)}
    grp_bank = np.zeros((ngroups, 16), dtype=np.int32)
    k = np.full(ne * N, -1, dtype=np.int64)
    xrank = np.full(nn, -1, dtype=np.int64)
    for v in np.argsort(first_slot, kind="stable"):
        L = int(cnt[v])
        sl = slot_lists[v]
        best_total, best_r, best_ks = None, None, None
        for r in range(16):
            if not free2[L][r]:
                continue
            avail = list(range(L))
            ks, total = [], 0
            for sidx in sl:
                g = gid[sidx]
                bb, bc = None, None
                for kk in avail:
                    b = (kk + r) % 16
                    c = grp_bank[g, b]
                    if bc is None or c < bc:
                        bb, bc = kk, c
                avail.remove(bb)
                ks.append(bb)
                total += bc
            if best_total is None or total < best_total:
                best_total, best_r, best_ks = total, r, ks
        xrank[v] = free2[L][best_r].pop(0)
        for sidx, kk in zip(sl, best_ks):
            k[sidx] = kk
            grp_bank[gid[sidx], (kk + best_r) % 16] += 1
    loc = rank[inv].reshape(ne, N)
    xcol = xrank[inv].reshape(ne, N)
    return loc, k.reshape(ne, N), lens_sorted, xcol


def main():
    import flat
    import dkt
    dim, md = 4, 10
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    xyz, lev = dkt.trees.moving_ball_tree(dim, level, md)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    tot = {"default": [0, 0, 0], "bank-aware": [0, 0, 0]}
    nchunks = 0
    for slots in sim.regular_chunks(t, md):
        nchunks += 1
        if nchunks > 60:
            break
        for name, fn in (("default", sim.default_assignment), ("bank-aware", bank_aware)):
            out = fn(slots)
            loc, k, lens = out[:3]
            pos = sim.positions(out[3] if len(out) > 3 else loc, k, lens)
            assert len(np.unique(pos)) == pos.size  # every slot owns one position
            wg, ni = sim.wavefronts(loc)
            ws, _ = sim.wavefronts(pos)
            tot[name][0] += wg
            tot[name][1] += ws
            tot[name][2] += ni
    for name, (g, s, ni) in tot.items():
        print("%-11s gather %.2f  scatter %.2f  wavefronts/instruction (ideal 2)" % (name, g / ni, s / ni))


if __name__ == "__main__":
    main()
