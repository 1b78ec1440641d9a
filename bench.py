#!/usr/bin/env python
"""Benchmark of the matvec hot path (BASELINE.json: "matvec DOF/s (fp64, 4D p=1 adaptive)").

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU matvec

One "step" = one matvec v = A u of the 4-D p=1 Laplacian on the class-B space-time moving-ball
tree.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dendro-kt_b200"))

METRIC = "matvec DOF/s (fp64, 4D p=1 adaptive)"
UNIT = "DOF/s"
DIM, ORDER, MAX_DEPTH = 4, 1, 12


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def kernel_source_sha():
    """sha256 (16 hex digits) over the CUDA sources of the matvec path: keys the ncu-measured traffic in profiles/."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "dendro-kt_b200", "csrc")
    for f in ("dkt_family.cu", "dkt_chunks.cu", "dkt_chunks.h", "dkt_dist.cu", "dkt_p2p.cuh"):
        with open(os.path.join(csrc, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic(n_elem):
    """DRAM bytes per matvec step from the committed ncu capture (profiles/traffic.json, written by tools/ncu_traffic.py):
    used only if it was taken with the same kernel sources on the same workload; None otherwise."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if t.get("kernel_source_sha") == kernel_source_sha() and int(t.get("n_elem", -1)) == int(n_elem):
            return float(t["dram_bytes_per_step"])
    except Exception:
        pass
    return None


GOLDEN_PARITY_CASES = ("ball-d4-p1-morton-5", "gauss-d4-p1-morton")


def golden_parity(dkt, rank, world, dist, bcast_id):
    """The bench's own DA path (partitioned when world > 1) on committed golden fixtures taken from the reference
    (tests/golden/*.npz, tests/golden/make_golden.py): v = A u for the unit-cell Laplacian (family kernel + singles) and a
    general dense operator (per-element kernels), gathered to the single-rank node order and compared with the reference's
    vectors.  Returns the worst relative error per case; the caller fails the run above 1e-12."""
    import numpy as np
    import torch
    from dkt import operators
    out = []
    for name in GOLDEN_PARITY_CASES:
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        dim, md = int(g["dim"]), int(g["max_depth"])
        kw = dict(ip0=g["ip0"], ip1=g["ip1"])
        if world > 1:
            kw.update(rank=rank, nranks=world, nccl_id=bcast_id())
        da = dkt.DA(g["in_xyz"], g["in_lev"], dim, 1, md, **kw)
        n = len(g["node_lev"])
        u = np.random.default_rng(99).uniform(-1.0, 1.0, n)  # the fixtures' input vector (tests/cases.py input_vector)
        worst = 0.0
        N = 1 << dim
        Kd = np.random.default_rng(5).uniform(-1.0, 1.0, (N, N))  # the fixtures' dense operator (tests/cases.py dense_operator)
        for op, scale, want in ((dkt.Operator.dense(operators.laplace_kref(dim, 1), dim - 2.0), 0.7, g["v_lap"]),
                                (dkt.Operator.dense(operators.laplace_kref(dim, 1), dim - 2.0, dirichlet=True), 0.7, g["v_lap_diri"]),
                                (dkt.Operator.dense(Kd, float(g["alpha"])), float(g["scale"]), g["v_dense"])):
            if world > 1:
                ids = torch.from_numpy(da.owned_ids().astype(np.int64)).cuda()
                v_loc = da.matvec(op, torch.from_numpy(u).cuda()[ids].contiguous(), scale=scale)
                torch.cuda.synchronize()  # the DA works on its own stream
                full = torch.zeros(n, dtype=torch.float64, device="cuda")
                full[ids] = v_loc
                dist.all_reduce(full)
                v = full.cpu().numpy()
            else:
                v = da.matvec(op, u, scale=scale)
            worst = max(worst, float(np.abs(v - want).max() / np.abs(want).max()))
        out.append({"case": name, "max_rel_err": worst})
        da.close()
    return out


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region: NVML polled every millisecond (nvidia_ml_py), `nvidia-smi` every
    0.1 s where NVML is not importable.  mark() brackets the timed region; the summary uses the samples inside it (all samples
    if the region was shorter than one polling interval)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.marks = index, [], False, []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def mark(self):
        self.marks.append(time.perf_counter())

    def _nvml_row(self):
        n = self.nvml
        sm = int(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        try:
            r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        bits = [0x8, 0x40, 0x20, 0x4]  # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap (nvml.h)
        return [str(sm), str(self.max_sm)] + ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag:
            t = time.perf_counter()
            try:
                if self.nvml is not None:
                    self.rows.append((t, self._nvml_row()))
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append((t, [c.strip() for c in out.split(",")]))
            except Exception:
                pass
            time.sleep(0.001 if self.nvml is not None else 0.1)

    def summary(self):
        rows = [r for _, r in self.rows]
        inside = [r for t, r in self.rows if len(self.marks) >= 2 and self.marks[0] <= t <= self.marks[-1]]
        use = inside if inside else rows
        sm = sorted(int(r[0]) for r in use if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({n for r in use if len(r) >= 6 for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "samples_in_timed_region": len(inside), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_reference_run(steps, warmup, level=5):
    """The reference's own CPU matvec (oracle/_ref = the reference compiled here; single MPI rank,
    single thread - the reference has no threading in this path) on a bounded sample of the
    workload: the same moving-ball tree at a smaller refinement level."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import dkt
    import dktref
    from dkt import operators
    if not dktref.available("morton"):
        raise RuntimeError("oracle/_ref is not built")
    xyz, lev = dkt.trees.moving_ball_tree(DIM, level, MAX_DEPTH)
    R = dktref.Reference(DIM, MAX_DEPTH)
    tree = R.tree_from_elements(xyz, lev, sort=True)
    da = R.da(tree, ORDER)
    K = operators.laplace_kref(DIM, ORDER)
    u = np.random.default_rng(99).uniform(-1, 1, da.num_nodes)
    _, secs, _ = da.matvec(u, dktref.OP_DENSE, K, alpha=DIM - 2.0, nwarm=warmup, niter=steps)
    sample = "4-D p=1 moving-ball tree at max_level %d (%d elements, %d nodes), %d matvecs after %d warm-up" % (
        level, len(lev), da.num_nodes, steps, warmup)
    return da.num_nodes / secs, secs, sample, da.num_nodes, len(lev)


def _reference_worker(idx, steps, warmup, level, barrier, out):
    """One single-rank reference process: builds the sample tree + ot::DA and times feMatrix::matVec."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import dkt
    import dktref
    from dkt import operators
    xyz, lev = dkt.trees.moving_ball_tree(DIM, level, MAX_DEPTH)
    R = dktref.Reference(DIM, MAX_DEPTH)
    tree = R.tree_from_elements(xyz, lev, sort=True)
    da = R.da(tree, ORDER)
    K = operators.laplace_kref(DIM, ORDER)
    u = np.random.default_rng(99 + idx).uniform(-1, 1, da.num_nodes)
    da.matvec(u, dktref.OP_DENSE, K, alpha=DIM - 2.0, nwarm=warmup, niter=0)
    barrier.wait()
    t0 = time.perf_counter()
    da.matvec(u, dktref.OP_DENSE, K, alpha=DIM - 2.0, nwarm=0, niter=steps)
    t1 = time.perf_counter()
    out.put((idx, t0, t1, da.num_nodes, len(lev)))


def _reference_mpi_job(rank, nranks, steps, warmup, level):
    """One rank of ONE distributed reference job (oracle/shim_mp): a contiguous piece of the sorted sample tree, the reference's
    distributed ot::DA, feMatrix::matVec with its ghost exchanges."""
    import numpy as np
    import dkt
    import dktref
    import dktref_mp
    from dkt import operators
    xyz, lev = dkt.trees.moving_ball_tree(DIM, level, MAX_DEPTH)
    R = dktref_mp.session(DIM, MAX_DEPTH)
    sx, sl = R.tree_from_elements(xyz, lev, sort=True).export()
    n = len(sl)
    lo, hi = rank * n // nranks, (rank + 1) * n // nranks
    da = R.da(R.tree_from_elements(sx[lo:hi], sl[lo:hi], sort=False), ORDER)
    info = dktref_mp.local_info(da)
    K = operators.laplace_kref(DIM, ORDER)
    u = np.random.default_rng(99 + rank).uniform(-1, 1, info[0])
    da.matvec(u, dktref.OP_DENSE, K, alpha=DIM - 2.0, nwarm=warmup, niter=0)
    t0 = time.perf_counter()
    da.matvec(u, dktref.OP_DENSE, K, alpha=DIM - 2.0, nwarm=0, niter=steps)  # every matVec synchronises the ranks (ghost exchange)
    return (time.perf_counter() - t0) / steps, info[0], info[2], info[5]


def reference_mpi_mode(args, procs, level):
    """The reference in its own parallel mode - ONE job, one rank per core, real ghost exchanges - over the multi-process MPI stand-in
    of oracle/shim_mp.  Reported beside the replica number (which stays the line's value: it is the upper bound)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import dktref_mp
        if not dktref_mp.available():
            return {"unavailable": "oracle/_ref/libdktref_mp_morton.so is not built"}
        ranks = max(1, min(procs, 16))
        res = dktref_mp.run(ranks, _reference_mpi_job, args.steps, args.warmup, level, arena_bytes=1 << 33, timeout=900)
        secs = max(r[0] for r in res)
        n_global = res[0][3]
        return {"ranks": ranks, "value": n_global / secs, "unit": UNIT, "ms_per_step": secs * 1e3, "n_nodes": n_global,
                "owned_nodes_min_max": [min(r[1] for r in res), max(r[1] for r in res)],
                "ghosted_nodes_min_max": [min(r[2] for r in res), max(r[2] for r in res)],
                "how": "one distributed job of the reference (its own partitioned ot::DA and ghost exchange) on the same sample tree, "
                       "ranks = processes over a shared-memory MPI stand-in (oracle/shim_mp); ONE tree shared by all cores, directly comparable with "
                       "`value` (the replicas' aggregate)"}
    except Exception as e:  # noqa: BLE001 - informational leg: never fails the arm
        return {"unavailable": repr(e)[:200]}


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores.  The reference parallelises by
    MPI rank per core and has no threading in this path; the image has no MPI, so every core runs an
    independent single-rank replica of the same sample (an upper bound for rank-per-core: no ghost
    exchange) and the aggregate rate is reported."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dkt
    import dktref
    import flat
    if not dktref.available("morton"):
        raise SystemExit("oracle/_ref is not built (python __graft_entry__.py in the build container)")
    dktref._lib("morton")  # mapped in THIS process too (the workers are forked from it)
    procs = args.ref_procs if args.ref_procs > 0 else max(1, min(os.cpu_count() or 1, 64))
    level = args.ref_level
    # workload facts of the sample, from the oracle's table construction (CPU, numpy)
    sx, sl = dkt.trees.moving_ball_tree(DIM, level, MAX_DEPTH)
    st = flat.build_tables(sx, sl, DIM, ORDER, MAX_DEPTH)
    n_hang, tree_class = int(len(st.hang_idx)), st.tree_class
    del sx, sl, st
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs)
    out = ctx.Queue()
    ws = [ctx.Process(target=_reference_worker, args=(i, args.steps, args.warmup, level, barrier, out)) for i in range(procs)]
    for w in ws:
        w.start()
    res = [out.get() for _ in ws]
    for w in ws:
        w.join()
    t0, t1 = min(r[1] for r in res), max(r[2] for r in res)
    n_nodes, n_elem = res[0][3], res[0][4]
    secs = (t1 - t0) / args.steps          # one step = one matvec on every replica
    value = procs * n_nodes / secs
    sample = ("%d independent single-rank replicas (one per core; no MPI in the image) of the 4-D p=1 moving-ball tree at max_level %d "
              "(%d elements, %d nodes each), %d matvecs after %d warm-up" % (procs, level, n_elem, n_nodes, args.steps, args.warmup))
    # The reference's real parallel mode is ONE job with a rank per core.  If that leg ran and is the faster of the two, it is the
    # arm's number (the reference at its best on this box); the replicas' aggregate stays in the line for comparison.
    replicas = {"value": value, "ms_per_step": secs * 1e3, "replicas": procs, "sample": sample}
    mpi = reference_mpi_mode(args, procs, level)
    mode = "replicas"
    if "value" in mpi and mpi["value"] > value:
        mode = "distributed"
        value, secs, procs = mpi["value"], mpi["ms_per_step"] * 1e-3, mpi["ranks"]
        sample = ("ONE distributed reference job, %d ranks (one per core) over the multi-process MPI stand-in oracle/shim_mp: the reference's own "
                  "partitioned ot::DA and ghost exchanges on the 4-D p=1 moving-ball tree at max_level %d (%d elements, %d nodes), %d matvecs "
                  "after %d warm-up" % (procs, level, n_elem, n_nodes, args.steps, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "4D p=1 Laplacian matvec (K_e = h^2 K_ref), space-time moving-ball adaptive tree, class B, max_level %d of "
                               "max_depth %d: the bench tree's generator at the largest level the reference builds and runs within about "
                               "a minute per core" % (level, MAX_DEPTH),
                   "dim": DIM, "order": ORDER, "n_elem": n_elem, "n_nodes": n_nodes, "n_hanging_elem": n_hang, "tree_class": tree_class,
                   "mode": mode, "cores": procs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_replicas": replicas,
        "reference_mpi": mpi,
    }
    print(json.dumps(line))


def weak_window(n_gpus, per_gpu_elems):
    """Time window of the moving ball such that the level-9 tree has ~per_gpu_elems * n_gpus
    elements (measured: elements ~= 9.4e6 + 3.41e8 * width at max_level 9)."""
    w = (n_gpus * per_gpu_elems - 9.4e6) / 3.41e8
    return min(max(w, 1.0 / 512), 0.5)


def run_gpu(args):
    import numpy as np
    import torch
    import dkt
    from dkt import operators
    if args.families is not None:
        os.environ["DKT_FAMILIES"] = str(args.families)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t)

    def bcast_id():
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(dkt.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return idt.cpu().numpy().tobytes()

    # ---- parity first: the same DA path on golden fixtures taken from the reference (all ranks take part) ----
    parity = golden_parity(dkt, rank, world, dist, bcast_id)
    if any(not (c["max_rel_err"] <= 1e-12) for c in parity):
        if rank == 0:
            print(json.dumps({"error": "parity check against the reference's golden vectors failed", "dist_parity": parity, "tol": 1e-12}))
        raise SystemExit(3)

    # ---- the tree: identical on every rank (weak scaling: the ball's time window grows with N) ----
    level = args.level
    width = weak_window(world, args.per_gpu_elems) if level == 9 else 0.5
    t0 = time.time()
    xyz, lev = dkt.trees.moving_ball_tree(DIM, level, MAX_DEPTH, use_torch=True, t0=0.5 - width / 2, t1=0.5 + width / 2)
    torch.cuda.synchronize()
    t_tree = time.time() - t0
    t0 = time.time()
    if world > 1:
        da = dkt.DA(xyz, lev, DIM, ORDER, MAX_DEPTH, rank=rank, nranks=world, nccl_id=bcast_id())
    else:
        da = dkt.DA(xyz, lev, DIM, ORDER, MAX_DEPTH)
    t_build = time.time() - t0
    n_elem_global = int(lev.numel())
    del xyz, lev
    torch.cuda.empty_cache()
    K = operators.laplace_kref(DIM, ORDER)
    op = dkt.Operator.dense(K, alpha=DIM - 2.0)
    n = da.n_nodes                      # owned by this rank
    n_global = da.n_global_nodes
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    da.set_stream(stream.cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(99 + rank)
    # device-resident vectors in the DA's ghosted layout [owned | ghosts] (as an application that keeps
    # its vectors on the GPU would hold them; the reference's matVec works on ghosted vectors too)
    ghosted = world > 1
    nloc = n + (da.n_ghost_nodes if ghosted else 0)
    u = torch.rand(nloc, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    v = torch.empty_like(u)

    # ---- device-resident throughput ("value") ---------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        da.matvec(op, u, v, ghosted=ghosted)
    barrier()
    launches0 = dkt.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    sampler.mark()
    ev[0].record(stream)
    for i in range(args.steps):
        da.matvec(op, u, v, ghosted=ghosted)
        ev[i + 1].record(stream)
    barrier()
    sampler.mark()
    sampler.stop_flag = True  # the end-to-end loop below is host-synchronous: no polling thread beside it
    sampler.join(timeout=2)
    launches = dkt.kernel_launch_count() - launches0
    my_ms = ev[0].elapsed_time(ev[-1]) / args.steps
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    per_rank = None
    if dist is not None:  # per-rank step time and work, to see partition imbalance
        t = torch.tensor([my_ms, float(da.n_mv_elem), float(da.n_hanging), float(da.n_nodes), float(da.n_ghost_nodes)],
                         dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [[round(float(x), 4) for x in r] for r in allt]
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    ms = total_ms / args.steps
    value = n_global / (ms * 1e-3)

    # ---- end to end through the host API: pinned host buffers, H2D + matvec + D2H every step ----
    uh = torch.empty(n, dtype=torch.float64).pin_memory()
    vh = torch.empty(n, dtype=torch.float64).pin_memory()
    uh.copy_(u[:n].cpu())
    un, vn = uh.numpy(), vh.numpy()
    for _ in range(min(args.warmup, 3)):
        da.matvec(op, un, vn)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        da.matvec(op, un, vn)  # synchronous: returns after the D2H copy has landed
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    check = float(np.abs(vn - v[:n].cpu().numpy()).max() / max(np.abs(vn).max(), 1e-300))

    peaks, peak_kind = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    alg_total = sum_over_ranks(float(da.alg_bytes))
    achieved = alg_total / world / (ms * 1e-3) / 1e9   # per-GPU average
    n_mv_total = sum_over_ranks(float(da.n_mv_elem))
    n_hang_total = sum_over_ranks(float(da.n_hanging))
    n_ghost_total = sum_over_ranks(float(da.n_ghost_nodes))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "4D p=1 Laplacian matvec (K_e = h^2 K_ref), space-time moving-ball adaptive tree, class B, "
                                   "max_level %d of max_depth %d, time window %.4f (grows with the GPU count: ~%.1e elements per GPU)"
                                   % (level, MAX_DEPTH, width, args.per_gpu_elems),
                       "dim": DIM, "order": ORDER, "n_elem": n_elem_global, "n_nodes": n_global, "n_visited_elem": int(n_mv_total),
                       "n_hanging_elem": int(n_hang_total), "n_ghost_nodes": int(n_ghost_total), "tree_class": da.tree_class,
                       "partition": "SFC-contiguous element ranges, one per GPU; NCCL send/recv ghost exchange" if world > 1 else "single GPU",
                       "cache": "working set %.0f MB per GPU > 126 MB L2 (no flush needed)" % (alg_total / world / 1e6),
                       "tables": "per-element (DKT_FAMILIES=0)" if os.environ.get("DKT_FAMILIES") == "0" else "sibling families + singles",
                       "tree_build_s": round(t_tree, 3), "da_build_s": round(t_build, 3), "chunks_rank0": da.chunk_info()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # ncu dram__bytes_read+write of the step's kernels on the default workload (profiles/README.md);
                         # None for other workloads
                         "traffic": measured_traffic(n_elem_global) if world == 1 else None,
                         "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write over one step, same kernel sources)",
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6650",
                         "alg_bytes_per_step_per_gpu": alg_total / world,
                         "kernel": "whole matvec step per GPU (memset + family kernel + per-element kernels of the singles" +
                                   (" + ghost pack/NCCL/unpack)" if world > 1 else ")")},
            "e2e": {"value": n_global / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * n_global, "d2h_bytes_per_step": 8 * n_global,
                    "ms_per_step": e2e_s * 1e3, "max_rel_diff_vs_device_path": check},
            "dist_parity": {"cases": parity, "tol": 1e-12, "ranks": world,
                            "what": "Laplacian (+ Dirichlet) and dense-operator matvec of this DA path on golden fixtures from the reference, "
                                    "gathered to the single-rank node order"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "ms_per_step_min_max_rank0": [min(per_step), max(per_step)],
            "per_rank_ms_elems_hanging_owned_ghost": per_rank,
        }
        if args.cpu_baseline and world == 1:
            try:
                cv, cs, sample, _, _ = cpu_reference_run(10, 10)
                line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample}
            except Exception as e:  # the bench line must survive a missing oracle
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "unavailable: %s" % e}
        print(json.dumps(line))
    da.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dkt", choices=["dkt", "reference"])
    ap.add_argument("--level", type=int, default=9, help="finest level of the moving-ball tree")
    ap.add_argument("--per-gpu-elems", type=float, default=1.2e7, help="weak scaling: target elements per GPU (level 9)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--ref-procs", type=int, default=0, help="--impl reference: replicas (0 = one per core, at most 64)")
    ap.add_argument("--ref-level", type=int, default=6, help="--impl reference: finest level of the sample tree (6: 6.5e5 elements, 36 %% hanging)")
    ap.add_argument("--families", type=int, default=None, choices=[0, 1],
                    help="0: per-element chunk tables only (DKT_FAMILIES=0); default: sibling-family tables")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
