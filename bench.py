#!/usr/bin/env python
"""bench.py — grid-cell·steps/s of the two-layer QG step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid M P]

A "step" is one pass of the hot path (evolve_zeta! + evolve_psi!, reference
src/run_model_no_output.jl:10-13) over the whole grid.  N = 1 runs BASELINE.json config 3, the
4096 x 4096 headline grid (dt = 5 min).  N > 1 (torchrun, one rank per
GPU) runs one independent 4096 x 4096 run per GPU — member-per-GPU ensemble sharding, no
data-path collective, weak scaling; timing is max over ranks.

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


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.ready = threading.Event()   # NVML is up (its start-up can outlast a short timed region)
        self.active = False              # samples count only while the timed region runs

    def begin(self):
        """Start the thread, wait until NVML answers, then count samples from here on."""
        self.start()
        self.ready.wait(timeout=20.0)
        self.active = True

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
                    time.sleep(0.001)
                    continue
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.005)
        except Exception as e:   # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_leg(M, P, steps, warmup, budget_s=25.0):
    """Times oracle/qg_oracle.c (the reference algorithm restated in C, OpenMP on all host
    threads) on the same grid for a bounded number of steps."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import qg_oracle as o
    import qg_oracle_c as oc
    a = model_args(M, P)
    m = o.make_model(*[a[k] for k in ("H_1", "H_2", "beta", "Lx", "Ly", "dt", "T", "U", "M", "P", "dx", "visc", "r",
                                      "R_d", "initial_kick")])
    zeta, psi = o.initialise_model(m, seed=1)
    f = np.zeros_like(zeta)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the C port sets
    # its thread count explicitly, so the launcher's default does not throttle the baseline)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else oc.max_threads()
    t = 1
    w = max(1, min(warmup, 1))
    t0 = time.perf_counter()
    oc.run_steps(m, zeta, psi, f, t, w, threads)
    per = (time.perf_counter() - t0) / w
    t += w
    n = int(max(1, min(steps, budget_s / max(per, 1e-9))))
    t0 = time.perf_counter()
    oc.run_steps(m, zeta, psi, f, t, n, threads)
    dt = time.perf_counter() - t0
    return {"value": M * P * n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} steps of the same {M}x{P} grid after {w} warm-up (C/OpenMP restatement, spectral "
                      f"solve standing in for CHOLMOD), {dt / n * 1e3:.1f} ms/step"}, dt / n * 1e3, n


def run_reference(args, rank, world):
    if rank != 0:
        return
    M, P = args.grid
    cb, ms, n = cpu_reference_leg(M, P, args.steps, args.warmup, budget_s=60.0)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"Phillips two-layer {M}x{P}, Float64, reference algorithm on host cores "
                                   f"(C restatement; Julia is not installed in this image)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_slab(args, rank, world, local_rank, torch, dist, qgb200, np):
    """BASELINE.json config 4: one run split into y-slabs over the GPUs of the node (NCCL halo
    ring + carry all-gather).  value = global cells x steps / max-over-ranks device time."""
    M, P = args.grid
    a = model_args(M, P)
    glob = qgb200.BaroclinicModel(*[a[k] for k in ("H_1", "H_2", "beta", "Lx", "Ly", "dt", "T", "U", "M", "P", "dx",
                                                   "visc", "r", "R_d", "initial_kick")])
    model = qgb200.slab.local_model(glob, world)
    K, W = args.steps, max(args.warmup, 3)
    # per-rank seeded slab: psi noise, q from the slab-periodic Laplacian (throughput does not depend on it)
    zeta, psi = qgb200.initialise_model(model, seed=1 + rank)
    f = np.zeros_like(zeta)
    ids = [qgb200.Session.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sess = qgb200.Session(model, members=1, device=local_rank, stream=stream.cuda_stream)
    sess.dist_init(rank, world, ids[0])

    def gather_blobs(blob):
        out = [None] * world
        dist.all_gather_object(out, blob)
        return out

    sess.dist_peer_init(gather_blobs)   # per-step exchanges over NVLink peer memory (QG_DIST_NCCL=1: stay on NCCL)
    sess.upload(zeta, psi, f)
    del zeta, psi, f
    sess.step(1, W)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.begin()
    sess.set_profiling(True)
    l0 = sess.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sess.step(W + 1, K)
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    sampler.active = False
    ms_total = e0.elapsed_time(e1)
    launches = sess.launch_count() - l0
    ktimes = sess.kernel_times()
    sess.set_profiling(False)
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    E, Z = sess.diagnostics()
    tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax[0])
    dist.barrier()     # nobody frees memory a peer still has mapped
    sess.close()
    if rank == 0:
        cells = float(M) * P
        value = cells * K / (ms_total * 1e-3)
        peak, peak_src = peaks()
        per = {k: (ms / n * 1e-3 if n else 0.0) for k, (ms, n) in ktimes.items()}
        local_cells = cells / world
        kern = {k: {"us": round(per[k] * 1e6, 2)} for k in per if per[k]}
        dom = "k1_zeta_step"
        ach = KERNEL_BYTES[dom] * local_cells / per[dom] / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"Phillips two-layer {M}x{P}, Float64, ONE run in {world} y-slabs of "
                                       f"{P // world} rows (BASELINE.json config 4): halo rows, k=0 column and y-solve "
                                       f"carries exchanged by "
                                       + ("NCCL send/recv + all-gather" if os.environ.get("QG_DIST_NCCL") else
                                          "in-kernel NVLink peer stores + flag barriers")
                                       + " in place of the all-to-all transpose",
                           "dt_s": a["dt"], "parallelism": f"y-slab x{world}",
                           "ic": "per-rank seeded slab (psi noise, q from the slab-periodic Laplacian)"},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                             "per_rank_kernels_us": kern,
                             "step": {"achieved_per_gpu": STEP_BYTES * local_cells * K / (ms_total * 1e-3) / 1e9,
                                      "frac": STEP_BYTES * local_cells * K / (ms_total * 1e-3) / 1e9 / peak}},
                "e2e": None, "gpu_launches": int(launches), "clocks": sampler.result(),
                "diagnostics": {"E": float(E), "Z": float(Z)}}
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
    ap.add_argument("--members", type=int, default=1,
                    help="ensemble members batched per GPU (BASELINE.json config 5: --grid 512 512 --members 8 on 8 GPUs)")
    ap.add_argument("--mode", default="ensemble", choices=["ensemble", "slab"],
                    help="N > 1: 'ensemble' = one independent run per GPU (default, weak scaling); "
                         "'slab' = ONE run of --grid split into y-slabs over the GPUs (strong scaling)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import qgb200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.mode == "slab" and world > 1:
        return run_slab(args, rank, world, local_rank, torch, dist, qgb200, np)
    M, P = args.grid
    a = model_args(M, P)
    model = qgb200.BaroclinicModel(*[a[k] for k in ("H_1", "H_2", "beta", "Lx", "Ly", "dt", "T", "U", "M", "P", "dx",
                                                    "visc", "r", "R_d", "initial_kick")])
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
    barrier()
    sampler.begin()
    # One timed region: K steps between two events on the launching stream, with a CUDA-event pair
    # around every kernel launch inside it (qg_set_profiling; the events are read back only after
    # the region), so the per-kernel durations of the roofline come from exactly these K steps.
    # Profiling uses plain launches; without it qg_step replays CUDA graphs of 3-step cycles
    # (worth < 1 % at 4096^2, 20-30 % on grids <= 512^2).
    sess.set_profiling(True)
    l0 = sess.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sess.step(W + 1, K)
    e1.record(stream)
    barrier()
    sampler.active = False
    ms_total = e0.elapsed_time(e1)
    launches = sess.launch_count() - l0
    ktimes = sess.kernel_times()
    sess.set_profiling(False)
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    E, Z = sess.diagnostics()
    E, Z = np.atleast_1d(E), np.atleast_1d(Z)
    if not (np.all(np.isfinite(E)) and np.all(np.isfinite(Z))):
        raise SystemExit("bench.py: state went non-finite during the timed region")

    # ---- the same K steps without the per-launch events (CUDA-graph replay), informational -------
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    sess.step(W + K + 1, K)
    g1.record(stream)
    barrier()
    ms_graph = g0.elapsed_time(g1)

    # ---- end-to-end timing through host buffers ("e2e") -------------------------------------
    views[2][...] = 0.0
    zeta0, psi0 = initial_state()
    views[0][...] = zeta0
    views[1][...] = psi0
    del zeta0, psi0
    Ke = K
    barrier()
    t0 = time.perf_counter()
    sess.upload_initial_raw(pin[0].data_ptr(), pin[1].data_ptr())   # level 1 of zeta, psi; f_store = 0
    sess.step(1, Ke)
    sess.download_raw(pin[0].data_ptr(), pin[1].data_ptr(), 0)      # all three levels of zeta, psi
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = 2 * (n_elem // 3) * 8
    d2h = 2 * n_elem * 8

    tmax = torch.tensor([ms_total, e2e_s * 1e3, ms_graph], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, ms_graph = float(tmax[0]), float(tmax[1]), float(tmax[2])
    sess.close()

    if rank == 0:
        cells = float(M) * P * mb      # per GPU
        value = world * cells * K / (ms_total * 1e-3)
        e2e_value = world * cells * Ke / (e2e_ms * 1e-3)
        peak, peak_src = peaks()
        per = {k: (ms / n * 1e-3 if n else 0.0) for k, (ms, n) in ktimes.items()}
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
            "config": {"workload": f"Phillips two-layer {M}x{P}, Float64, "
                                   + (f"ensemble of {world * mb} members, {mb} batched per B200 (BASELINE.json config 5 "
                                      f"shape; member sharding, no collective)" if mb > 1 else
                                      "single B200 (BASELINE.json config 3, headline roofline run)" if world == 1 else
                                      f"{world} independent members, one per B200 (ensemble sharding, no collective)"),
                       "dt_s": a["dt"], "ic": "seeded uniform psi noise on shear U (initialise_model), seed 1+rank",
                       "l2": (f"working set {18 * (M + 18) * (P + 4) * 8 * mb / 1e9 + 16 * M * P * mb / 1e9:.2f} GB per GPU "
                              f"vs 126 MB L2, no explicit flush"),
                       "parallelism": f"member-per-gpu x{world}" + (f" ({mb} members per GPU)" if mb > 1 else "")},
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
                    "ms_total": e2e_ms},
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
            "diagnostics": {"E": float(E[0]), "Z": float(Z[0])},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_reference_leg(M, P, 4, 1, budget_s=20.0)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
