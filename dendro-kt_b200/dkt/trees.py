"""Seeded generators of 2:1-balanced linear trees (inputs of the matvec path).

The reference builds its benchmark trees on the host from point clouds
(SFC_Tree::distTreeBalancing, src/tsort.cpp:862-878; ~60 us/element in 4-D) - that is outside
the matvec path and is NOT re-implemented here.  These generators produce trees that are
2:1-balanced (all 3^dim-1 neighbours) BY CONSTRUCTION from a refinement criterion:

    split cell C  <=>  g(centre(C)) < K * h(C),  g 2-Lipschitz in the max-norm, K = 3

(for a leaf A of size h touching a finer subdivided cell P of size h/2: |cA-cP|_inf <= 3h/4, so
g(cA) <= g(cP) + 3h/2 < K h/2 + 3h/2 <= K h => A would have been split; hence neighbouring
leaves differ by at most one level).  Everything is vectorised level by level and runs either in
NumPy (tests) or on the GPU in torch (benchmarks).
"""
import numpy as np


def uniform_tree(dim, level, max_depth):
    """Regular grid 2^(dim*level) (Example2 of test/testAdaptiveExamples.h, ot::DA regular ctor)."""
    n = 1 << level
    ax = np.arange(n, dtype=np.uint32) << np.uint32(max_depth - level)
    g = np.stack(np.meshgrid(*([ax] * dim), indexing="ij"), axis=-1).reshape(-1, dim)
    return np.ascontiguousarray(g), np.full(len(g), level, dtype=np.uint8)


def uniform_tree_torch(dim, level, max_depth, device="cuda"):
    import torch
    n = 1 << level
    ax = torch.arange(n, dtype=torch.int32, device=device) << (max_depth - level)
    g = torch.stack(torch.meshgrid(*([ax] * dim), indexing="ij"), dim=-1).reshape(-1, dim).contiguous()
    return g, torch.full((g.shape[0],), level, dtype=torch.uint8, device=device)


def _refine(xp, dim, max_depth, min_level, max_level, gfun, K=3.0, device=None):
    """Level-by-level refinement; xp is numpy or torch."""
    is_torch = xp.__name__ == "torch"
    if is_torch:
        cells = xp.zeros((1, dim), dtype=xp.int64, device=device)
    else:
        cells = xp.zeros((1, dim), dtype=xp.int64)
    out_xyz, out_lev = [], []
    nch = 1 << dim
    if is_torch:
        offs = xp.tensor([[(c >> d) & 1 for d in range(dim)] for c in range(nch)], dtype=xp.int64, device=device)
    else:
        offs = xp.array([[(c >> d) & 1 for d in range(dim)] for c in range(nch)], dtype=xp.int64)
    for lvl in range(0, max_level + 1):
        h = 1.0 / (1 << lvl)
        size = 1 << (max_depth - lvl)
        ctr = (cells.to(xp.float64) if is_torch else cells.astype(xp.float64)) / float(1 << max_depth) + 0.5 * h
        if lvl < min_level:
            split = xp.ones(cells.shape[0], dtype=xp.bool_) if not is_torch else xp.ones(cells.shape[0], dtype=xp.bool, device=device)
        elif lvl >= max_level:
            split = xp.zeros(cells.shape[0], dtype=xp.bool_) if not is_torch else xp.zeros(cells.shape[0], dtype=xp.bool, device=device)
        else:
            split = gfun(ctr) < K * h
        leaf = cells[~split]
        out_xyz.append(leaf)
        if is_torch:
            out_lev.append(xp.full((leaf.shape[0],), lvl, dtype=xp.uint8, device=device))
        else:
            out_lev.append(xp.full((leaf.shape[0],), lvl, dtype=xp.uint8))
        par = cells[split]
        if par.shape[0] == 0:
            break
        cells = (par[:, None, :] + offs[None, :, :] * (size >> 1)).reshape(-1, dim)
    if is_torch:
        return xp.cat(out_xyz).to(xp.int32).contiguous(), xp.cat(out_lev).contiguous()
    return xp.concatenate(out_xyz).astype(xp.uint32), xp.concatenate(out_lev)


def _ball_g(xp, dim, radius, c0, vel, t0, t1):
    """Distance-like function (2-Lipschitz in max-norm for dim<=4, |vel|<=0.25) to the surface of
    a ball of `radius` whose centre moves along axis 1 with the last coordinate (time)."""
    def g(ctr):
        if dim == 4:
            t = ctr[:, 3]
            tc = xp.clip(t, t0, t1) if xp.__name__ == "numpy" else xp.clamp(t, t0, t1)
            dx = ctr[:, 0] - c0[0]
            dy = ctr[:, 1] - (c0[1] + vel * tc)
            dz = ctr[:, 2] - c0[2]
            rho = xp.sqrt(dx * dx + dy * dy + dz * dz)
            a = xp.abs(rho - radius)
            b = xp.abs(t - tc)
            return xp.maximum(a, b) if xp.__name__ == "numpy" else xp.maximum(a, b)
        d2 = 0.0
        for d in range(dim):
            dd = ctr[:, d] - c0[d]
            d2 = d2 + dd * dd
        return xp.abs(xp.sqrt(d2) - radius)
    return g


def moving_ball_tree(dim, max_level, max_depth, min_level=4, radius=0.125, use_torch=False, device="cuda", t0=0.25, t1=0.75):
    """The class-B space-time moving-ball tree (after test/testMovingBall.cpp:
    122-175): a sphere of radius 0.125 centred at (0.375, 0.375+0.25 t, 0.375) for t in
    [t0, t1] (default [0.25, 0.75]; keep it inside (0.2, 0.8)), refined to `max_level` at its surface, uniform level-`min_level` (>= 4) guard at the domain
    boundary so that no level jump touches it.  For dim < 4 a static ball at (0.375,...)."""
    c0 = (0.375, 0.375, 0.375)
    if use_torch:
        import torch
        return _refine(torch, dim, max_depth, min_level, max_level, _ball_g(torch, dim, radius, c0, 0.25, t0, t1), device=device)
    return _refine(np, dim, max_depth, min_level, max_level, _ball_g(np, dim, radius, c0, 0.25, t0, t1))


def gaussian_points(dim, n, max_depth, seed=7, sigma=0.04, guard_level=None):
    """The reference benchmark's point distribution (include/octUtils.h:29-66: normal, mean 0.5,
    sigma 1/25, clamped) with a fixed seed; optional level-`guard_level` cell centres (class B)."""
    rng = np.random.default_rng(seed)
    pts = np.clip(rng.normal(0.5, sigma, (n, dim)), 0.0, 1.0 - 1e-12)
    pts = (pts * (1 << max_depth)).astype(np.uint32)
    if guard_level is not None:
        m = 1 << guard_level
        g = np.stack(np.meshgrid(*([np.arange(m)] * dim), indexing="ij"), -1).reshape(-1, dim)
        gp = ((g + 0.5) / m * (1 << max_depth)).astype(np.uint32)
        pts = np.concatenate([pts, gp])
    return pts


def guard_points(dim, guard_level, max_depth):
    """Centres of all level-`guard_level` cells: a uniform guard so that every boundary-touching leaf has the same level."""
    m = 1 << guard_level
    g = np.stack(np.meshgrid(*([np.arange(m)] * dim), indexing="ij"), -1).reshape(-1, dim)
    return ((g + 0.5) / m * (1 << max_depth)).astype(np.uint32)


def shell_points(dim, n, max_depth, seed=1331, radius=0.125, guard_level=3):
    """points on a sphere shell moving through time (after test/testMovingBall.cpp:122-175, moved off the
    domain boundary), plus a level-`guard_level` guard.  4-D: (x, y, z, t); lower dim drops trailing space axes and keeps t."""
    rng = np.random.default_rng(seed)
    theta = rng.uniform(0.0, 2.0 * np.pi, n)
    zu = rng.uniform(-1.0, 1.0, n)
    tau = rng.uniform(0.0, 1.0, n)
    r, h = np.sqrt(np.abs(zu)), np.sqrt(1.0 - np.abs(zu))
    t = 0.25 + 0.5 * tau
    sp = np.stack([0.375 + radius * r * np.cos(theta), 0.375 + 0.25 * t + radius * r * np.sin(theta), 0.375 + radius * np.sign(zu) * h], 1)
    p = np.concatenate([sp[:, :dim - 1], t[:, None]], 1)
    pts = (np.clip(p, 0.0, 1.0 - 1e-12) * (1 << max_depth)).astype(np.uint32)
    if guard_level is not None:
        pts = np.concatenate([pts, guard_points(dim, guard_level, max_depth)])
    return pts
