#!/usr/bin/env python
"""bench.py — grid-cell·steps/s of the two-layer QG step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid M P]

A "step" is one pass of the hot path (evolve_zeta! + evolve_psi!, reference
src/run_model_no_output.jl:10-13) over the whole grid.

N = 1 runs BASELINE.json config 3, the 4096 x 4096 headline grid (dt = 5 min), and adds
`config4_single_gpu`: the 16384 x 8192 grid of config 4 on the one GPU (the strong-scaling baseline).

N > 1 (torchrun, one rank per GPU) prints ONE line with two things in it:
  * top level (`value`, `e2e`, `roofline`): one independent 4096 x 4096 run per GPU — member-per-GPU
    ensemble sharding, no data-path collective, weak scaling, max over ranks;
  * `slab`: BASELINE.json config 4 — ONE 16384 x 8192 run split into N y-slabs (halo rows, k = 0 column
    and y-solve carries exchanged by in-kernel NVLink peer stores) — strong scaling: `slab.parity`
    (a 2048 x 4096 slab run of 10 steps against the C oracle's global solution), `slab.ms_per_step`,
    per-rank kernel times, clocks, and in the same run rank 0's single-GPU time of the same grid and
    steps (`slab.single_gpu`), whose energy / enstrophy the slab run must reproduce.

value  : cell·steps/s with the state resident in HBM, K steps between CUDA events.
e2e    : the same metric through the public host API on pinned HOST buffers: upload of the
         reference-layout state arrays, K steps, download of (zeta, psi) — the
         run_model_no_output call pattern — all inside the timed region.
roofline: the dominant kernel's algorithmic bytes / its CUDA-event duration measured inside
         the timed region, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference: the C restatement of the reference algorithm
         (oracle/qg_oracle.c, OpenMP, all host threads) on a bounded sample of the same grid.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))

METRIC = "grid-cell·steps/s, Phillips 2-layer QG at 4096²"
UNIT = "cell·steps/s"
# algorithmic bytes per cell per launch (DESIGN.md "Roofline accounting")
KERNEL_BYTES = {"k1_zeta_step": 96.0, "k2_fft_forward": 32.0, "k3_ysolve": 32.0, "k4_fft_inverse": 32.0}
STEP_BYTES = 128.0   # implementation-independent compulsory traffic per cell·step
MODEL_KEYS = ("H_1", "H_2", "beta", "Lx", "Ly", "dt", "T", "U", "M", "P", "dx", "visc", "r", "R_d", "initial_kick")
CONFIG4 = (16384, 8192)
SLAB_PARITY_GRID = (2048, 4096)


def model_args(M, P):
    """The reference's parameter block (src/benchmarking/benchmarking.jl:6-18) on the
    benchmark grids; dt per SURVEY.md 8d (stable explicit viscosity)."""
    KM, MIN = 1000.0, 60.0
    Lx = 4000.0 * KM
    dx = Lx / M
    Ly = dx * P
    dt = 60.0 * MIN if M <= 1024 else (5.0 * MIN if M <= 4096 else 30.0)
    return dict(H_1=1.0 * KM, H_2=2.0 * KM, beta=2e-11, Lx=Lx, Ly=Ly, dt=dt, T=86400.0, U=0.1, M=M, P=P, dx=dx,
                visc=100.0, r=1e-7, R_d=40.0 * KM, initial_kick=1e-6)


def baseline_config_name(M, P):
    return {(128, 128): "BASELINE.json config 1 grid", (1024, 1024): "BASELINE.json config 2",
            (4096, 4096): "BASELINE.json config 3, headline roofline run",
            (16384, 8192): "BASELINE.json config 4 grid on one GPU", (512, 512): "BASELINE.json config 5 grid"}.get(
                (M, P), "custom grid")


def workload_config(M, P, world, mb, dt):
    """`config` of the JSON line; the reference arm prints the same dictionary."""
    if mb > 1:
        what = (f"ensemble of {world * mb} members, {mb} batched per B200 (BASELINE.json config 5 shape; member "
                f"sharding, no collective)")
    elif world == 1:
        what = f"single B200 ({baseline_config_name(M, P)})"
    else:
        what = f"{world} independent members, one per B200 (ensemble sharding, no collective)"
    return {"workload": f"Phillips two-layer {M}x{P}, Float64, {what}", "dt_s": dt,
            "ic": "seeded uniform psi noise on shear U (initialise_model), seed 1+rank",
            "l2": (f"working set {18 * (M + 18) * (P + 4) * 8 * mb / 1e9 + 16 * M * P * mb / 1e9:.2f} GB per GPU "
                   f"vs 126 MB L2, no explicit flush"),
            "parallelism": f"member-per-gpu x{world}" + (f" ({mb} members per GPU)" if mb > 1 else "")}


def ncu_traffic(kernel, M, P):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if list(t["grid"]) != [M, P]:
            return None
        return t["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def bind_near_gpu(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE the pinned host buffers
    are allocated, so they land on that NUMA node (8 ranks staging through one socket's memory was
    the e2e limiter).  Returns (previous affinity, description)."""
    if not hasattr(os, "sched_getaffinity"):
        return None, "no sched_setaffinity"
    prev = os.sched_getaffinity(0)
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 64
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        near = cpus & prev
        if near and near != prev:
            os.sched_setaffinity(0, near)
            return prev, f"bound to {len(near)} of {len(prev)} CPUs local to GPU {index}"
        return prev, "GPU-local CPU set equals the process affinity (single NUMA node or already bound)"
    except Exception as e:
        return prev, f"not bound ({type(e).__name__})"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.ready = threading.Event()   # NVML is up (its start-up can outlast a short timed region)
        self.active = False              # samples count only while a timed region runs

    def begin(self):
        """Start the thread (once), wait until NVML answers, then count samples from here on."""
        if not self.is_alive() and not self.ready.is_set():
            self.start()
            self.ready.wait(timeout=20.0)
        self.active = True

    def pause(self):
        self.active = False

    def finish(self):
        self.active = False
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=2.0)
        return self.result()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                     "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.ready.set()
            while not self.stop_flag:
                if not self.active:
                    time.sleep(0.0005)
                    continue
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.001)
        except Exception as e:   # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores (oracle/ is the checker and the CPU baseline only)
# ---------------------------------------------------------------------------------------------------
def oracle_modules():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qg_oracle as o
    import qg_oracle_c as oc
    return o, oc


def host_threads(oc):
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the C port sets
    # its thread count explicitly, so the launcher's default does not throttle the baseline)
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else oc.max_threads()


def cpu_reference_leg(M, P, steps, warmup, budget_s=25.0):
    """Times oracle/qg_oracle.c (the reference algorithm restated in C, OpenMP on all host
    threads) on the same grid: `warmup` untimed steps, then `steps` timed ones, both cut to what
    fits `budget_s` seconds (the cut is stated in `sample`)."""
    import numpy as np
    o, oc = oracle_modules()
    a = model_args(M, P)
    m = o.make_model(*[a[k] for k in MODEL_KEYS])
    zeta, psi = o.initialise_model(m, seed=1)
    f = np.zeros_like(zeta)
    threads = host_threads(oc)
    t0 = time.perf_counter()
    oc.run_steps(m, zeta, psi, f, 1, 1, threads)
    per = time.perf_counter() - t0
    w = int(max(1, min(warmup, 0.25 * budget_s / max(per, 1e-9))))
    if w > 1:
        oc.run_steps(m, zeta, psi, f, 2, w - 1, threads)
    n = int(max(1, min(steps, 0.75 * budget_s / max(per, 1e-9))))
    t0 = time.perf_counter()
    oc.run_steps(m, zeta, psi, f, 1 + w, n, threads)
    dt = time.perf_counter() - t0
    cut = "" if (n == steps and w == warmup) else f" (asked for {steps} after {warmup}; cut to the {budget_s:.0f} s budget)"
    return {"value": M * P * n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} steps of the same {M}x{P} grid after {w} warm-up{cut} (C/OpenMP restatement, spectral "
                      f"solve standing in for CHOLMOD), {dt / n * 1e3:.1f} ms/step"}, dt / n * 1e3, n, w


def run_reference(args, rank, world):
    if rank != 0:
        return
    M, P = args.grid
    cb, ms, n, w = cpu_reference_leg(M, P, args.steps, args.warmup, budget_s=120.0)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": w, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(M, P, max(1, args.gpus), max(1, args.members), model_args(M, P)["dt"]),
            "impl_note": "reference algorithm on the host cores of rank 0: oracle/qg_oracle.c, the C/OpenMP restatement "
                         "(Julia is not installed in this image, so the reference's own CHOLMOD path cannot run)",
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def make_model(qgb200, M, P):
    a = model_args(M, P)
    return qgb200.BaroclinicModel(*[a[k] for k in MODEL_KEYS]), a


def per_launch_seconds(ktimes):
    return {k: (ms / n * 1e-3 if n else 0.0) for k, (ms, n) in ktimes.items()}


def timed_steps(torch, stream, sess, first, K, barrier, sampler=None, profile=True):
    """K steps between two CUDA events on the launching stream (optionally with a CUDA-event pair
    around every kernel launch inside the region).  Returns (ms, launches, per-kernel times)."""
    barrier()
    if sampler is not None:
        sampler.begin()
    if profile:
        sess.set_profiling(True)
    l0 = sess.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sess.step(first, K)
    e1.record(stream)
    barrier()
    if sampler is not None:
        sampler.pause()
    ms = e0.elapsed_time(e1)
    launches = sess.launch_count() - l0
    ktimes = sess.kernel_times() if profile else {}
    if profile:
        sess.set_profiling(False)
    return ms, launches, ktimes


def single_gpu_run(torch, qgb200, np, stream, device, M, P, W, K, sampler=None):
    """One GPU, device-resident, device-drawn initial condition (seed 1): W warm-up + K timed steps."""
    model, a = make_model(qgb200, M, P)
    with qgb200.Session(model, members=1, device=device, stream=stream.cuda_stream) as s:
        s.init_state(1)
        s.step(1, W)
        ms, launches, kt = timed_steps(torch, stream, s, W + 1, K, torch.cuda.synchronize, sampler)
        E, Z = s.diagnostics()
    per = per_launch_seconds(kt)
    peak, _ = peaks()
    cells = float(M) * P
    return {"grid": [M, P], "dt_s": a["dt"], "steps": K, "warmup": W, "ms_per_step": ms / K,
            "value": cells * K / (ms * 1e-3), "unit": UNIT,
            "kernels_us": {k: round(v * 1e6, 2) for k, v in per.items() if v},
            "kernels_frac": {k: round(KERNEL_BYTES[k] * cells / per[k] / 1e9 / peak, 4) for k in KERNEL_BYTES if per.get(k)},
            "step_frac_128B": STEP_BYTES * cells * K / (ms * 1e-3) / 1e9 / peak,
            "gpu_launches": int(launches), "E": float(E), "Z": float(Z),
            "ic": "device initial condition, seed 1 (qg_init_state)"}


def slab_session(torch, dist, qgb200, stream, local_rank, rank, world, M, P):
    """One rank of a y-slab run of the global M x P grid; exchanges over NVLink peer memory."""
    glob, a = make_model(qgb200, M, P)
    model = qgb200.slab.local_model(glob, world)
    ids = [qgb200.Session.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sess = qgb200.Session(model, members=1, device=local_rank, stream=stream.cuda_stream)
    sess.dist_init(rank, world, ids[0])

    def gather_blobs(blob):
        out = [None] * world
        dist.all_gather_object(out, blob)
        return out

    peer = True
    if not os.environ.get("QG_DIST_NCCL"):
        peer = sess.dist_peer_init(gather_blobs)
    else:
        peer = False
    return sess, glob, a, peer


def slab_parity_leg(torch, dist, qgb200, np, stream, local_rank, rank, world, steps=10, grid=None):
    """A y-slab run of SLAB_PARITY_GRID against the C oracle's solution of the GLOBAL problem.
    Rank 0 draws the same device initial condition on a single-GPU handle, downloads it, lets the
    oracle step it on the host, and broadcasts the result; every rank compares its slab."""
    M, P = grid or SLAB_PARITY_GRID
    o, oc = oracle_modules()
    sess, glob, a, peer = slab_session(torch, dist, qgb200, stream, local_rank, rank, world, M, P)
    shape = (M + 2, P + 2, 2)
    ref = torch.empty((2,) + shape[::-1], dtype=torch.float64, device="cuda")   # oracle (zeta, psi) level 1, C order = F order reversed
    ez = torch.zeros(2, dtype=torch.float64, device="cuda")
    t_oracle = 0.0
    if rank == 0:
        with qgb200.Session(glob, members=1, device=local_rank, stream=stream.cuda_stream) as s1:
            s1.init_state(1)
            zeta, psi = s1.new_state_array(), s1.new_state_array()
            s1.download(zeta=zeta, psi=psi)
        mo = o.make_model(*[a[k] for k in MODEL_KEYS])
        f = np.zeros_like(zeta)
        t0 = time.perf_counter()
        oc.run_steps(mo, zeta, psi, f, 1, steps, host_threads(oc))
        t_oracle = time.perf_counter() - t0
        Eo, Zo = o.diagnostics(mo, zeta, psi)
        ref[0].copy_(torch.from_numpy(np.ascontiguousarray(zeta[:, :, :, 0].T)))
        ref[1].copy_(torch.from_numpy(np.ascontiguousarray(psi[:, :, :, 0].T)))
        ez[0], ez[1] = float(Eo), float(Zo)
        del zeta, psi, f
    dist.broadcast(ref, src=0)
    dist.broadcast(ez, src=0)
    sess.init_state(1)
    sess.step(1, steps)
    zl, pl = sess.new_state_array(), sess.new_state_array()
    sess.download(zeta=zl, psi=pl)
    E, Z = sess.diagnostics()
    dist.barrier()     # nobody frees memory a peer still has mapped
    sess.close()
    pl_rows = P // world
    j0 = rank * pl_rows
    refh = ref.cpu().numpy()   # [2][layer][P+2][M+2]
    rel = lambda x, y: float(np.abs(x - y).max() / np.abs(y).max())
    errs, where = [], []
    for k, mine in ((0, zl), (1, pl)):
        want = refh[k][:, j0:j0 + pl_rows + 2, :]                        # ghost-inclusive rows of this slab
        got = np.ascontiguousarray(mine[:, :, :, 0].T)                    # [layer][P_loc+2][M+2]
        errs.append(max(rel(got[l], want[l]) for l in range(2)))
        d = np.abs(got - want)
        l, jj, ii = np.unravel_index(int(np.argmax(d)), d.shape)
        where.append({"rank": rank, "layer": int(l), "local_row": int(jj) - 1, "col": int(ii) - 1,
                      "interior_only": max(rel(got[x][1:-1, 1:-1], want[x][1:-1, 1:-1]) for x in range(2))})
    Eo, Zo = float(ez[0]), float(ez[1])
    errs += [abs(E - Eo) / abs(Eo), abs(Z - Zo) / abs(Zo)]
    allerrs, allwhere = [None] * world, [None] * world
    dist.all_gather_object(allerrs, errs)
    dist.all_gather_object(allwhere, where)
    worst = np.max(np.array(allerrs, dtype=np.float64), axis=0)
    worst_rank = int(np.argmax(np.array(allerrs, dtype=np.float64)[:, 1]))
    ok = bool(worst[0] <= 1e-10 and worst[1] <= 1e-10 and worst[2] <= 1e-8 and worst[3] <= 1e-8)
    return {"grid": [M, P], "steps": steps, "ranks": world, "q": float(worst[0]), "psi": float(worst[1]),
            "E": float(worst[2]), "Z": float(worst[3]), "ok": ok,
            "tolerance": "q, psi <= 1e-10 relative (max norm, per layer, ghost rows included), E, Z <= 1e-8",
            "against": f"oracle/qg_oracle.c on the global grid from the same device-drawn initial condition "
                       f"({t_oracle:.1f} s on the host)",
            "exchange": "NVLink peer stores + flag barriers" if peer else "NCCL",
            "worst_cell": {"q": allwhere[int(np.argmax(np.array(allerrs, dtype=np.float64)[:, 0]))][0],
                           "psi": allwhere[worst_rank][1]},
            "per_rank": [[float(x) for x in e[:2]] for e in allerrs]}


def slab_timed_leg(torch, dist, qgb200, np, stream, local_rank, rank, world, M, P, W, K):
    """BASELINE.json config 4: ONE run of M x P in `world` y-slabs; W warm-up + K timed steps from the
    device initial condition (seed 1, identical to the single-GPU field)."""
    sess, glob, a, peer = slab_session(torch, dist, qgb200, stream, local_rank, rank, world, M, P)
    sess.init_state(1)
    sess.step(1, W)

    def barrier():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    ms, launches, kt = timed_steps(torch, stream, sess, W + 1, K, barrier, sampler)
    clocks = sampler.finish()
    E, Z = sess.diagnostics()
    gaps_us = (ms - sum(t for (t, n) in kt.values())) / K * 1e3   # this rank: flag barriers + launch gaps per step
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax[0])
    dist.barrier()     # nobody frees memory a peer still has mapped
    sess.close()
    per = per_launch_seconds(kt)
    allper, allclk = [None] * world, [None] * world
    dist.all_gather_object(allper, {k: round(v * 1e6, 2) for k, v in per.items() if v})
    dist.all_gather_object(allclk, clocks)
    cells = float(M) * P
    peak, _ = peaks()
    kmax = {k: max(p.get(k, 0.0) for p in allper) for k in allper[0]}
    return {"grid": [M, P], "dt_s": a["dt"], "ranks": world, "rows_per_rank": P // world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "value": cells * K / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
            "exchange": ("in-kernel NVLink peer stores + flag barriers" if peer else "NCCL send/recv + all-gather")
                        + " in place of the all-to-all transpose",
            "per_rank_kernels_us_max": kmax, "per_rank_kernels_us_rank0": allper[0],
            "barriers_and_gaps_us_rank0": round(gaps_us, 1),
            "step_frac_128B_per_gpu": STEP_BYTES * cells / world * K / (ms * 1e-3) / 1e9 / peak,
            "gpu_launches_rank0": int(launches), "clocks_per_rank": allclk,
            "E": float(E), "Z": float(Z), "ic": "device initial condition, seed 1 (qg_init_state, global indexing)"}


def slab_legs(torch, dist, qgb200, np, stream, local_rank, rank, world, W, K):
    """Everything the N > 1 line reports about config 4.  Every rank returns the same dictionary."""
    out = {"config": "BASELINE.json config 4: Phillips two-layer 16384x8192, y-slab decomposition"}
    out["parity"] = slab_parity_leg(torch, dist, qgb200, np, stream, local_rank, rank, world)
    M, P = CONFIG4
    out.update(slab_timed_leg(torch, dist, qgb200, np, stream, local_rank, rank, world, M, P, W, K))
    # the same grid, initial condition and steps on rank 0's GPU alone: the strong-scaling baseline
    single = [None]
    if rank == 0:
        single[0] = single_gpu_run(torch, qgb200, np, stream, local_rank, M, P, W, K)
    dist.broadcast_object_list(single, src=0)
    s1 = single[0]
    out["single_gpu"] = {k: s1[k] for k in ("ms_per_step", "value", "kernels_us", "E", "Z")}
    out["speedup_vs_single_gpu"] = s1["ms_per_step"] / out["ms_per_step"]
    dE, dZ = abs(out["E"] - s1["E"]) / abs(s1["E"]), abs(out["Z"] - s1["Z"]) / abs(s1["Z"])
    out["full_size_check"] = {"what": f"energy / enstrophy after {W + K} steps of 16384x8192: {world}-slab run vs the "
                                      f"single-GPU run from the same initial condition", "E_rel": dE, "Z_rel": dZ,
                              "ok": bool(dE <= 1e-8 and dZ <= 1e-8)}
    return out


def run_slab_only(args, rank, world, local_rank, torch, dist, qgb200, np):
    """Developer mode (--mode slab): only the y-slab run of --grid, printed as the line itself."""
    M, P = args.grid
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    r = slab_timed_leg(torch, dist, qgb200, np, stream, local_rank, rank, world, M, P, max(args.warmup, 3), args.steps)
    if rank == 0:
        peak, peak_src = peaks()
        cells = float(M) * P / world
        k1 = r["per_rank_kernels_us_max"].get("k1_zeta_step", 0.0) * 1e-6
        ach = KERNEL_BYTES["k1_zeta_step"] * cells / k1 / 1e9 if k1 else 0.0
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": r["steps"],
                "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"Phillips two-layer {M}x{P}, Float64, ONE run in {world} y-slabs of "
                                       f"{P // world} rows (BASELINE.json config 4): {r['exchange']}",
                           "dt_s": r["dt_s"], "parallelism": f"y-slab x{world}", "ic": r["ic"]},
                "roofline": {"bound": "hbm", "kernel": "k1_zeta_step", "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak, "traffic": None, "peak_source": peak_src},
                "slab": r, "e2e": None, "gpu_launches": r["gpu_launches_rank0"],
                "clocks": r["clocks_per_rank"][0]}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--grid", type=int, nargs=2, default=[4096, 4096])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the 16384x8192 legs (config4_single_gpu / slab)")
    ap.add_argument("--config4-timeout", type=int, default=300)
    ap.add_argument("--members", type=int, default=1,
                    help="ensemble members batched per GPU (BASELINE.json config 5: --grid 512 512 --members 8 on 8 GPUs)")
    ap.add_argument("--mode", default="ensemble", choices=["ensemble", "slab"],
                    help="N > 1: 'ensemble' = the contract line (one independent run per GPU at top level + the "
                         "config-4 y-slab legs under `slab`); 'slab' = ONLY one y-slab run of --grid (developer mode)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    prev_affinity, numa_note = bind_near_gpu(local_rank)

    import numpy as np
    import torch
    import qgb200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=600))
    if args.mode == "slab" and world > 1:
        return run_slab_only(args, rank, world, local_rank, torch, dist, qgb200, np)
    M, P = args.grid
    model, a = make_model(qgb200, M, P)
    K, W = args.steps, args.warmup
    # synthetic "randomly perturbed jet": seeded white-noise psi on the uniform shear U (rank = member)
    mb = max(1, args.members)

    def initial_state():
        if mb == 1:
            return qgb200.initialise_model(model, seed=1 + rank)
        zs, ps = zip(*[qgb200.initialise_model(model, seed=1 + rank * mb + m) for m in range(mb)])
        return np.asfortranarray(np.stack(zs, axis=-1)), np.asfortranarray(np.stack(ps, axis=-1))

    zeta, psi = initial_state()
    n_elem = zeta.size
    pin = [torch.empty(n_elem, dtype=torch.float64).pin_memory() for _ in range(3)]
    views = [p.numpy().reshape(zeta.shape, order="F") for p in pin]
    views[0][...] = zeta
    views[1][...] = psi
    views[2][...] = 0.0
    del zeta, psi

    # a dedicated (non-default) stream shared by torch's events and the library's launches
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sess = qgb200.Session(model, members=mb, device=local_rank, stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") --------------------------------------------------
    sess.upload_raw(pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr())
    sess.step(1, W)
    sampler = ClockSampler(local_rank)
    # One timed region: K steps between two events on the launching stream, with a CUDA-event pair
    # around every kernel launch inside it (qg_set_profiling; the events are read back only after
    # the region), so the per-kernel durations of the roofline come from exactly these K steps.
    # Profiling uses plain launches; without it qg_step replays CUDA graphs of 3-step cycles
    # (worth < 1 % at 4096^2, 20-30 % on grids <= 512^2).
    ms_total, launches, ktimes = timed_steps(torch, stream, sess, W + 1, K, barrier, sampler)
    E, Z = sess.diagnostics()
    E, Z = np.atleast_1d(E), np.atleast_1d(Z)
    if not (np.all(np.isfinite(E)) and np.all(np.isfinite(Z))):
        raise SystemExit("bench.py: state went non-finite during the timed region")

    # ---- the same K steps without the per-launch events (CUDA-graph replay), informational -------
    ms_graph, _, _ = timed_steps(torch, stream, sess, W + K + 1, K, barrier, sampler, profile=False)

    # ---- end-to-end timing through host buffers ("e2e") -------------------------------------
    views[2][...] = 0.0
    zeta0, psi0 = initial_state()
    views[0][...] = zeta0
    views[1][...] = psi0
    del zeta0, psi0
    Ke = K
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    sess.upload_initial_raw(pin[0].data_ptr(), pin[1].data_ptr())   # level 1 of zeta, psi; f_store = 0
    sess.step(1, Ke)
    sess.download_raw(pin[0].data_ptr(), pin[1].data_ptr(), 0)      # all three levels of zeta, psi
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.finish()
    clocks["sampled_over"] = "the three timed regions (per-kernel events, graph replay, e2e)"
    h2d = 2 * (n_elem // 3) * 8
    d2h = 2 * n_elem * 8

    tmax = torch.tensor([ms_total, e2e_s * 1e3, ms_graph], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, ms_graph = float(tmax[0]), float(tmax[1]), float(tmax[2])
    sess.close()
    del pin, views

    line = None
    if rank == 0:
        cells = float(M) * P * mb      # per GPU
        value = world * cells * K / (ms_total * 1e-3)
        e2e_value = world * cells * Ke / (e2e_ms * 1e-3)
        peak, peak_src = peaks()
        per = per_launch_seconds(ktimes)
        dom = max(KERNEL_BYTES, key=lambda k: per.get(k, 0.0))
        ach = KERNEL_BYTES[dom] * cells / per[dom] / 1e9
        kern = {k: {"us": round(per[k] * 1e6, 2), "GBps": round(KERNEL_BYTES[k] * cells / per[k] / 1e9, 1),
                    "frac": round(KERNEL_BYTES[k] * cells / per[k] / 1e9 / peak, 4)}
                for k in KERNEL_BYTES if per.get(k)}
        small = {k: round(per[k] * 1e6, 2) for k in per if k not in KERNEL_BYTES and per[k]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "ms_per_step_graph_replay": ms_graph / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(M, P, world, mb, a["dt"]),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": ncu_traffic(dom, M, P), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": KERNEL_BYTES[dom] * cells,
                         "algorithmic_bytes_per_cell": KERNEL_BYTES[dom],
                         "step": {"achieved": STEP_BYTES * cells * K / (ms_total * 1e-3) / 1e9,
                                  "frac": STEP_BYTES * cells * K / (ms_total * 1e-3) / 1e9 / peak,
                                  "algorithmic_bytes_per_cell_step": STEP_BYTES},
                         "kernels": kern, "small_kernels_us": small,
                         "kernel_timing": "CUDA-event pair around every launch inside the timed region"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / Ke, "d2h_bytes_per_step": d2h / Ke,
                    "what": f"run_model_no_output call pattern on pinned host arrays: qg_upload_initial_state(zeta, psi) -> "
                            f"qg_step({Ke}) -> qg_download_state(zeta, psi: 3 levels)",
                    "ms_total": e2e_ms, "host_numa": numa_note},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "diagnostics": {"E": float(E[0]), "Z": float(Z[0])},
        }

    # ---- BASELINE.json config 4 (16384 x 8192): single GPU at N = 1, y-slabs at N > 1 ---------------------
    if not args.no_config4 and mb == 1 and [M, P] == [4096, 4096]:
        # The headline line must survive the extra legs: if they have not finished in time (a rank died
        # inside a collective), rank 0 prints the line without them and every rank leaves.
        def give_up():
            if rank == 0:
                line["slab" if world > 1 else "config4_single_gpu"] = {"error": f"not finished after {args.config4_timeout} s"}
                print(json.dumps(line), flush=True)
            os._exit(0)
        watchdog = threading.Timer(args.config4_timeout + (0 if rank == 0 else 5), give_up)
        watchdog.daemon = True
        watchdog.start()
        try:
            if world == 1:
                extra = single_gpu_run(torch, qgb200, np, stream, local_rank, CONFIG4[0], CONFIG4[1], W, K)
                extra["what"] = "BASELINE.json config 4 grid on ONE B200: the baseline of the y-slab strong scaling"
                if line is not None:
                    line["config4_single_gpu"] = extra
            else:
                extra = slab_legs(torch, dist, qgb200, np, stream, local_rank, rank, world, W, K)
                if line is not None:
                    line["slab"] = extra
        except Exception as e:   # the headline line must survive a failure of the extra legs
            if line is not None:
                line["slab" if world > 1 else "config4_single_gpu"] = {"error": f"{type(e).__name__}: {e}"}
            print(f"bench.py: config-4 leg failed on rank {rank}: {type(e).__name__}: {e}", file=sys.stderr, flush=True)
        watchdog.cancel()

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            if prev_affinity is not None:
                os.sched_setaffinity(0, prev_affinity)   # the CPU baseline uses every host thread again
            cb, _, _, _ = cpu_reference_leg(M, P, 4, 1, budget_s=20.0)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if dist is not None:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
