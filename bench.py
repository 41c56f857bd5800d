#!/usr/bin/env python
"""bench.py — registrations/sec of the SAC-COT hot path (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload cfg2_3dmatch_256x5000|cfg3_...|cfg4_...|cfg5_single_n50000] [--scaling weak|strong]

Default workload = BASELINE.json configs[1] (3DMatch-scale synthetic batch: 256 pairs x N=5000 correspondences,
5 % inliers, tau_c = 0.1 m).  One "step" = one pass of the whole hot path (graph -> triangle counts -> COT selection
-> 3-point Kabsch -> K x N scoring + argmax -> fp64 refit) over one batch of synthetic pairs.

Batched workloads (cfg2/3/4): pairs are independent, so multi-GPU = one process per GPU (torchrun), no data-path
collective.  --scaling weak (default; what the driver's scaling run uses): every GPU gets its own 256 pairs.
--scaling strong: the SAME 256-pair batch, pair b on GPU b mod G.
cfg5 (single pair, N = 50 000): sac_cot_register_sharded — triangle cells and hypothesis ranges split over the ranks,
ncclAllGather + ncclAllReduce(max) inside the library; strong scaling by construction.

  value : whole-job registrations/s with the inputs already resident in HBM (SAC_COT_LOC_DEVICE entry points on
          torch's current stream), timed with CUDA events per step, L2 flushed between steps, max over ranks.
  e2e   : the same metric through the host-buffer C-ABI call (pinned host inputs, H2D and D2H inside the timed region).
  roofline     : the dominant kernel (triangle counting), timed live with CUDA events inside the library during
                 the timed steps, against the binding pipe's peak measured in this same process.
  cpu_baseline : the from-paper oracle on this box's host cores, rank 0, N = 1 only.

--impl reference times the oracle (OpenMP build, all host threads) on a bounded sample of the same workload: the
upstream repository ships no code (/root/reference/README.md:1-2), so the from-paper CPU oracle is the only
"reference implementation" there is.
"""
import argparse
import ctypes
import datetime
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sac_cot_b200 import _abi, synth  # noqa: E402
from sac_cot_b200.api import Registrar, load_library  # noqa: E402

WORKLOAD = "cfg2_3dmatch_256x5000"   # BASELINE.json configs[1]; --workload selects another config
UNIT = "registrations/s"
STAGES = ("pack", "graph", "scan", "theta", "triangles", "triangles_kept", "select", "apex", "kabsch", "score", "finalize",
          "exchange1", "exchange2")
SM_COUNT = 148
FP32_LANES = 128   # FP32 lanes per SM


def metric_name(cfg):
    return f"registrations/sec at N={cfg['N']} corr, {ratio_label(cfg)} inliers"


def ratio_label(cfg):
    return "/".join(f"{100 * r:g}%" for r in cfg["ratios"])


def make_batch(indices, seed_shift=0):
    """Pairs `indices` of the workload's batch, packed back to back."""
    cfg = synth.CONFIGS[WORKLOAD]
    ps = [synth.make_config_pair(WORKLOAD, b, seed_shift=seed_shift) for b in indices]
    src = np.ascontiguousarray(np.concatenate([p.src for p in ps]))
    dst = np.ascontiguousarray(np.concatenate([p.dst for p in ps]))
    offsets = np.arange(len(ps) + 1, dtype=np.int64) * cfg["N"]
    return ps, src, dst, offsets


def load_oracle(omp):
    name = "libsaccot_oracle_omp.so" if omp else "libsaccot_oracle.so"
    path = os.path.join(ROOT, "oracle", name)
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), name], check=True, stdout=subprocess.DEVNULL)
    return _abi.bind(ctypes.CDLL(path))


def time_oracle(lib, ps, cfg, threads=None):
    """Seconds per pair of the oracle over the given pairs."""
    with Registrar(lib=lib, tau_compat=cfg["tau"], tau_inlier=cfg["tau"]) as reg:
        if threads:
            reg.set("threads", threads)   # a launcher (torchrun) may have exported OMP_NUM_THREADS=1
        t0 = time.perf_counter()
        res = reg.register_batch([p.src for p in ps], [p.dst for p in ps])
        dt = time.perf_counter() - t0
        threads = reg.get("threads")
    return dt / len(ps), threads, res


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  The sampling process is started when the object is
    made (before the warm-up: nvidia-smi needs a few hundred ms before its first line, longer than a short timed
    region); start() / stop() mark the timed region and only the samples whose timestamps fall inside it are used."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t_begin = self.t_end = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def start(self):
        self.t_begin = time.time()

    @staticmethod
    def _when(stamp):
        try:   # "2026/10/18 23:11:02.123", local time
            return datetime.datetime.strptime(stamp, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        self.t_end = time.time()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            self.f.close()
            os.unlink(self.f.name)
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = []   # (time or None, sm, smax, power, reasons)
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                row = (self._when(c[0]), float(c[2]), float(c[3]), float(c[4]),
                       {name for name, v in zip(names, c[6:10]) if v.lower().startswith("active")})
            except ValueError:
                continue
            rows.append(row)
        self.f.close()
        os.unlink(self.f.name)
        t0 = self.t_begin if self.t_begin is not None else 0.0
        inside = [r for r in rows if r[0] is None or t0 - 0.02 <= r[0] <= self.t_end + 0.02]
        window = "timed region"
        if not inside and rows:
            # a region shorter than the sampling period: the samples of the second before it (the warm-up steps,
            # the same kernels on the same device)
            inside = [r for r in rows if r[0] is not None and t0 - 1.0 <= r[0] <= self.t_end + 0.02] or rows[-3:]
            window = "no sample fell inside the timed region; taken from the warm-up steps just before it"
        if inside:
            sm = [r[1] for r in inside]
            power = [r[3] for r in inside]
            reasons = set().union(*[r[4] for r in inside])
            # "under load" = samples at or above the median power draw
            pm = statistics.median(power)
            load = [s for s, p in zip(sm, power) if p >= pm] or sm
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(r[2] for r in inside), power_w_max=max(power),
                       reasons=sorted(reasons), samples=len(inside), window=window)
        return out


def run_reference(args, rank, world, emit):
    """--impl reference: the from-paper oracle (all host threads) on a bounded sample per step.  A CPU arm: rank 0
    alone runs it, with every host core whatever the launcher exported, so its value does not depend on --gpus."""
    if rank != 0:
        return
    cfg = synth.CONFIGS[WORKLOAD]
    lib = load_oracle(omp=True)
    cores = os.cpu_count() or 1
    single = cfg["pairs"] == 1
    sample = 1 if single else 32
    steps = 1 if single else args.steps
    ps, _, _, _ = make_batch(range(sample))
    if not single:
        for _ in range(args.warmup):
            time_oracle(lib, ps[:8], cfg, cores)
    t0 = time.perf_counter()
    threads = 1
    for _ in range(steps):
        _, threads, _ = time_oracle(lib, ps, cfg, cores)
    dt = time.perf_counter() - t0
    value = steps * sample / dt
    what = (f"the single N={cfg['N']} pair, registered once (no warm-up; --steps ignored: one registration takes "
            "tens of seconds on the host)" if single else
            f"bounded sample: the first {sample} of the {cfg['pairs']} pairs per step")
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 0 if single else args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "strong" if single or args.scaling == "strong" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: N={cfg['N']}, {ratio_label(cfg)} inliers, tau_c={cfg['tau']}",
                   "step": what, "sample_pairs_per_step": sample, "K_e": 1024, "apex_per_edge": 4,
                   "host_threads": int(threads),
                   "note": "CPU arm: independent of --gpus (rank 0 runs it on every host core)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(threads), "kind": "port",
                         "sample": f"{sample} pair(s)/step x {steps} step(s), from-paper oracle, OpenMP build "
                                   "(upstream repo has no code to run)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def fp32_lane_rate(sm_mhz):
    """FP32 instruction-lane issue peak of the device, lanes/s (148 SMs x 128 lanes x clock)."""
    return SM_COUNT * FP32_LANES * sm_mhz * 1e6


def stage_report(reg, steps):
    us = {s: reg.get(f"stage_us_{s}") for s in STAGES}
    calls = {s: reg.get(f"stage_calls_{s}") for s in STAGES}
    total = max(1, sum(us.values()))
    return us, calls, {
        "stage_share": {s: us[s] / total for s in STAGES if calls[s]},
        "stage_us_per_step": {s: us[s] / steps for s in STAGES if calls[s]},
    }


def tensor_roofline(reg, N, tri_us, units_per_launch, flop_share, sm_mhz, peaks, traffic):
    """Roofline object of the tensor-core triangle kernel.  units_per_launch pairs (x flop_share of each: 1/world
    when sharded) per launch of tri_us microseconds."""
    npad = (N + 127) // 128 * 128
    flop_pair = float(N) * (N - 1) * N            # N(N-1)/2 node pairs x N MACs x 2 flop
    tiles_pair = sum(min((N + 255) // 256, (240 * jq + 238) // 256 + 1) for jq in range((N + 239) // 240))
    kpad = (npad + 511) // 512 * 512
    exec_flop_pair = tiles_pair * 256.0 * 240.0 * kpad * 2.0
    achieved = flop_pair * flop_share * units_per_launch / (tri_us * 1e-6) / 1e12
    executed = exec_flop_pair * flop_share * units_per_launch / (tri_us * 1e-6) / 1e12
    peak = reg.get("probe_mxf4_gflops") / 1e3      # TFLOP/s, measured now on this GPU
    bf16 = peaks.get("bf16_tflops_sustained")
    return {
        "kernel": "triangles_mma_kernel (S2, tcgen05 kind::mxf4 cta_group::2, operands expanded on chip)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic,
        "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/ncu_traffic.json)",
        "peak_source": "measured in this process: the kernel's own MMA (mxf4 block-scaled, cta_group::2, M256 N240 K64, "
                       "same shared-memory operand layout) issued back to back on every CTA pair with nothing else "
                       "running (library probe `probe_mxf4_gflops`; MEASURED_PEAKS.json has no 4-bit entry)",
        "executed_frac": executed / peak,
        "algorithmic_flop_per_pair": flop_pair, "executed_flop_per_pair": exec_flop_pair,
        "algorithmic_flop_per_launch": flop_pair * flop_share * units_per_launch,
        "note_bf16": {"bf16_tflops_sustained": bf16, "achieved_over_bf16": achieved / bf16 if bf16 else None,
                      "why": "for reference only: 4-bit MMAs are not bounded by the bf16 pipe rate (4x per clock)"},
        "launch_us": tri_us, "pairs_per_launch": units_per_launch,
    }, flop_pair / (peak * 1e12) * 1e6


def cpu_baseline_batched(pairs, ps, cfg, inliers_gpu):
    n1, nall = min(pairs, 24), min(pairs, 64)
    s1, _, res1 = time_oracle(load_oracle(omp=False), ps[:n1], cfg)
    sall, threads, _ = time_oracle(load_oracle(omp=True), ps[:nall], cfg, os.cpu_count())
    agree = bool((res1.inliers == inliers_gpu[:n1]).all())
    return {
        "value": 1.0 / s1, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"first {n1} of the {pairs} pairs, from-paper C++ oracle, 1 thread ({s1:.3f} s/pair); "
                  "the upstream repo has no code to run",
        "all_cores": {"value": 1.0 / sall, "cores": int(threads), "sample": f"first {nall} pairs, OpenMP build"},
        "host_cpus": os.cpu_count(), "inlier_counts_match_gpu": agree,
    }


def run_batched(args, rank, local_rank, world, emit, torch, dist, dev):
    cfg = synth.CONFIGS[WORKLOAD]
    total_pairs = args.pairs or cfg["pairs"]
    N = cfg["N"]
    strong = args.scaling == "strong"
    if strong:   # the same batch, pair b on GPU b mod G
        mine = list(range(rank, total_pairs, world))
        ps, src, dst, offsets = make_batch(mine)
        job_pairs = total_pairs
    else:        # every GPU its own batch
        ps, src, dst, offsets = make_batch(range(total_pairs), seed_shift=rank * cfg["pairs"])
        job_pairs = world * total_pairs
    pairs = len(ps)

    lib = load_library()
    stream = torch.cuda.Stream(dev)  # every kernel/copy of the library and every timing event lives on this stream
    torch.cuda.set_stream(stream)
    reg = Registrar(lib=lib, device=local_rank, stream=stream.cuda_stream, tau_compat=cfg["tau"],
                    tau_inlier=cfg["tau"])
    if args.chunk_pairs:
        reg.set("chunk_pairs", args.chunk_pairs)
    if args.lanes:
        reg.set("lanes", args.lanes)
    if args.triangle_path >= 0:
        reg.set("triangle_path", args.triangle_path)
    if args.triangle_dbg:
        reg.set("triangle_dbg", args.triangle_dbg)
    if args.tile_runs >= 0:
        reg.set("tile_runs", args.tile_runs)
    if args.node_prune >= 0:
        reg.set("node_prune", args.node_prune)
    K = reg.params.num_edges * reg.params.apex_per_edge

    # device-resident inputs / outputs
    d_src = torch.from_numpy(src).to(dev)
    d_dst = torch.from_numpy(dst).to(dev)
    d_R = torch.empty((pairs, 3, 3), dtype=torch.float32, device=dev)
    d_t = torch.empty((pairs, 3), dtype=torch.float32, device=dev)
    d_i = torch.empty(pairs, dtype=torch.int32, device=dev)
    # pinned host inputs / outputs for the end-to-end leg
    h_src = torch.from_numpy(src).pin_memory()
    h_dst = torch.from_numpy(dst).pin_memory()
    h_R = torch.empty((pairs, 3, 3), dtype=torch.float32).pin_memory()
    h_t = torch.empty((pairs, 3), dtype=torch.float32).pin_memory()
    h_i = torch.empty(pairs, dtype=torch.int32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        reg.register_packed_ptr(d_src.data_ptr(), d_dst.data_ptr(), offsets, d_R.data_ptr(), d_t.data_ptr(),
                                d_i.data_ptr(), _abi.LOC_DEVICE)

    def step_host():
        reg.register_packed_ptr(h_src.data_ptr(), h_dst.data_ptr(), offsets, h_R.data_ptr(), h_t.data_ptr(),
                                h_i.data_ptr(), _abi.LOC_HOST)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up (also grows the workspace / key pool to steady state) ----
    sampler = ClockSampler(local_rank)   # sampling starts now; start() / stop() bracket the timed region
    for _ in range(args.warmup):
        step_device()
        torch.cuda.synchronize(dev)
        reg.get("last_status")   # a key-pool growth voids that step; the next one is clean
    step_device()
    torch.cuda.synchronize(dev)
    assert reg.get("last_status") == 0, "workspace did not reach steady state during warm-up"
    # sanity: the device-resident results recover the ground-truth poses
    R_chk, t_chk = d_R.cpu().numpy(), d_t.cpu().numpy()
    ok = 0
    for b in range(pairs):
        ang, dt_ = synth.pose_error(R_chk[b], t_chk[b], ps[b].R_gt, ps[b].t_gt)
        ok += ang < np.deg2rad(5.0) and dt_ < 1.5 * cfg["tau"]
    recall = ok / pairs

    # ---- timed region 1: device-resident (value) ----
    launches0 = reg.get("launches")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # evict L2 between steps (not timed)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    barrier()
    wall_dev = time.perf_counter() - wall0
    clocks = sampler.stop()
    assert reg.get("last_status") == 0
    launches = reg.get("launches") - launches0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(ms_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = job_pairs * args.steps / (total_ms * 1e-3)

    # ---- per-kernel pass: the same K steps on ONE lane (no chunk overlap), every stage bracketed by
    # CUDA events on the stream it runs on.  With several lanes the stages of different chunks overlap, so their
    # event spans measure contention, not kernel time. ----
    lanes_default = reg.get("lanes")
    reg.set("lanes", 1)
    step_device()
    torch.cuda.synchronize(dev)
    reg.set("stage_timing", 1)
    ev1 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev1[0].record(stream)
    for k in range(args.steps):
        step_device()
    ev1[1].record(stream)
    torch.cuda.synchronize(dev)
    serial_ms_per_step = ev1[0].elapsed_time(ev1[1]) / args.steps
    stage_us, stage_calls, stage_common = stage_report(reg, args.steps)
    reg.set("stage_timing", 0)
    reg.set("lanes", lanes_default)
    # exact node pruning of S2 (kernels_prune.cu) in the last chunk of the pass above
    node_prune = {"mode": reg.get("node_prune"), "pruned_pairs_last_chunk": reg.get("pruned_pairs"),
                  "kept_nodes_last_chunk": reg.get("kept_nodes"), "trying": reg.get("node_prune_trying"),
                  "what": "pairs whose selectable edges provably join few high-degree nodes count triangles for those "
                          "nodes' rows only (exact; --node-prune 0 switches it off).  trying = 0: the ctx saw two calls "
                          "in a row prune nothing and skips the attempt for its next node_prune_probe (30) calls"}

    # ---- timed region 2: end to end through the host-buffer C-ABI call (e2e) ----
    step_host()  # warm the host path (arena regrows once: it now also holds the input copy)
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()  # H2D (pinned) -> pipeline -> D2H -> stream sync, all inside the call
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = job_pairs * args.steps / float(e2e_s.item())
    # host-path results equal the device-resident ones
    same = bool((h_R.numpy() == R_chk).all() and (h_t.numpy() == t_chk).all())

    # ---- single-pair latency through sac_cot_register's path (one pair per call, host buffers; p50 of 100) ----
    latency = None
    if rank == 0 and not args.no_latency:
        latency = {}
        for name, n_pts in (("cfg1_n1000", 1000), ("n5000", 5000)):
            q = synth.make_pair(n_pts, 0.10 if n_pts == 1000 else 0.05, 42)
            hs, hd = torch.from_numpy(q.src).pin_memory(), torch.from_numpy(q.dst).pin_memory()
            off1 = np.array([0, n_pts], dtype=np.int64)
            oR, ot, oi = (torch.empty(9).pin_memory(), torch.empty(3).pin_memory(),
                          torch.empty(1, dtype=torch.int32).pin_memory())
            ts = []
            for k in range(110):
                t0 = time.perf_counter()
                reg.register_packed_ptr(hs.data_ptr(), hd.data_ptr(), off1, oR.data_ptr(), ot.data_ptr(), oi.data_ptr(),
                                        _abi.LOC_HOST)
                ts.append(time.perf_counter() - t0)
            ts = sorted(ts[10:])
            latency[name] = {"p50_us": 1e6 * ts[len(ts) // 2], "p90_us": 1e6 * ts[int(len(ts) * 0.9)],
                             "min_us": 1e6 * ts[0], "calls": len(ts), "inliers": int(oi[0])}
        latency["what"] = ("wall clock of one host-buffer single-pair call (pinned buffers; H2D, 11 kernels, D2H, one "
                           "stream sync) on the bench ctx")

    # ---- roofline of the dominant kernel (triangle counting), from the live stage timers ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    path_used = reg.get("triangle_path_used")  # which S2 kernels the timed steps ran
    reg.set("keep_debug", 1)
    small = min(pairs, 8)
    reg.set("triangle_path", 0)  # every edge key is kept on this path: exact edge count per pair
    reg.register_packed(src[: small * N], dst[: small * N], offsets[: small + 1])
    E_mean = float(np.mean([int(reg.debug(b, _abi.DBG_NUM_EDGES)[0]) for b in range(small)]))
    reg.set("keep_debug", 0)
    reg.set("triangle_path", args.triangle_path if args.triangle_path >= 0 else 2)
    npad = (N + 127) // 128 * 128
    stride = npad // 32
    tri_calls = max(1, stage_calls["triangles"])
    tri_us = stage_us["triangles"] / tri_calls                 # average launch duration (CUDA events, live)
    pairs_per_launch = pairs * args.steps / tri_calls
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tj_all = json.load(open(tpath)) if os.path.exists(tpath) and WORKLOAD == "cfg2_3dmatch_256x5000" else {}
    timing_note = (f"CUDA events around every stage over {args.steps} steps on one lane "
                   f"({serial_ms_per_step:.2f} ms/step without chunk overlap; the headline uses {lanes_default} lanes)")
    lane_rate = fp32_lane_rate(sm_mhz)
    P = N * (N - 1) / 2.0
    floors = {"graph_us": 21.0 * P / lane_rate * 1e6,          # 21 individually rounded fp32 ops per node pair
              "score_us": 16.0 * K * N / lane_rate * 1e6}      # 16 FP32 instructions per (hypothesis, point)
    node_pruned = path_used == 1 and node_prune["pruned_pairs_last_chunk"] > 0
    if node_pruned:
        # S2 ran on the kept rows only (exact node pruning): the dominant kernel is the compatibility graph.  It is
        # bound by FP32 issue (21 individually rounded fp32 operations per node pair, packed two per instruction
        # where the order of operations allows), neither by HBM nor by the tensor cores; both views are given.
        g_calls = max(1, stage_calls["graph"])
        g_us = stage_us["graph"] / g_calls
        g_pairs = pairs * args.steps / g_calls
        peak_gbs = peaks.get("hbm_gbs") or 6650.0
        g_bytes_pair = npad * stride * 4 * 2 + 24 * npad        # adjacency + its K-panel copy written, SoA points read
        ops = 21.0 * P * g_pairs
        roofline = {
            "kernel": "graph_kernel (S1, compatibility graph; S2 runs on the kept rows only)", "bound": "fp32",
            "achieved": ops / (g_us * 1e-6) / 1e12, "peak": lane_rate / 1e12, "unit": "Tops/s (fp32 lane operations)",
            "frac": ops / (g_us * 1e-6) / lane_rate, "traffic": None,
            "peak_source": f"148 SMs x 128 FP32 lanes x {sm_mhz:.0f} MHz (this run's clock under load)",
            "algorithmic_ops_per_launch": ops, "launch_us": g_us, "pairs_per_launch": g_pairs,
            "hbm": {"algorithmic_bytes_per_launch": g_bytes_pair * g_pairs,
                    "achieved_gbs": g_bytes_pair * g_pairs / (g_us * 1e-6) / 1e9, "peak_gbs": peak_gbs,
                    "frac": g_bytes_pair * g_pairs / (g_us * 1e-6) / 1e9 / peak_gbs},
            "note": "the contract's bound is hbm | tensor; this kernel is bound by neither (arithmetic intensity ~ 40 fp32 "
                    "operations per byte written), so the binding resource is reported and the HBM view given beside it",
            "edges_per_pair": E_mean,
            "triangles_kept_us_per_step": stage_us["triangles_kept"] / args.steps,
        }
        floors["triangles_us"] = 0.0   # the kept rows' share of S2 is not modelled (data dependent)
    elif path_used == 1:
        traffic = None
        if "triangles_mma_kernel" in tj_all:
            tj = tj_all["triangles_mma_kernel"]
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["pairs_in_launch"] * pairs_per_launch
        roofline, tri_floor_us = tensor_roofline(reg, N, tri_us, pairs_per_launch, 1.0, sm_mhz, peaks, traffic)
        roofline.update(edges_per_pair=E_mean)
        floors["triangles_us"] = tri_floor_us
    else:
        peak_gbs = peaks.get("hbm_gbs") or 6650.0
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)"
        # algorithmic bytes per pair of S2: read A once, write one 8-byte key per edge,
        # read-modify-write the per-node sums and the histogram once
        tri_bytes_pair = npad * stride * 4 + E_mean * 8 + npad * 8 * 2 + 4096 * 4 * 2
        achieved_gbs = tri_bytes_pair * pairs_per_launch / (tri_us * 1e-6) / 1e9
        wordops = E_mean * stride * pairs_per_launch               # AND+POPC on 32-bit words
        popc_peak = 16 * SM_COUNT * sm_mhz * 1e6                   # 16 POPC/clk/SM (15.9 measured, profiles/pipes_r01.txt)
        traffic = None
        if "triangles_block_kernel" in tj_all:
            tj = tj_all["triangles_block_kernel"]
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["pairs_in_launch"] * pairs_per_launch
        roofline = {
            "kernel": "triangles_block_kernel (S2, POPC bitset)", "bound": "hbm", "achieved": achieved_gbs,
            "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs, "traffic": traffic,
            "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/ncu_traffic.json)",
            "algorithmic_bytes_per_launch": tri_bytes_pair * pairs_per_launch, "peak_source": peak_src,
            "algorithmic_bytes_per_pair": tri_bytes_pair,
            "note": "S2 is bound by POPC/ALU issue, not HBM (arithmetic intensity ~ 40 word-ops/B): the HBM "
                    "fraction is small by construction; the issue-side figure is in `issue`",
            "issue": {"wordops_per_s": wordops / (tri_us * 1e-6), "popc_peak_per_s": popc_peak,
                      "frac": wordops / (tri_us * 1e-6) / popc_peak, "edges_per_pair": E_mean, "words_per_row": stride},
            "launch_us": tri_us, "pairs_per_launch": pairs_per_launch,
        }
        # 3 POPC per 5 words (carry-save adders) is what the kernel issues at best
        floors["triangles_us"] = E_mean * stride * 0.6 / popc_peak * 1e6
    us_per_pair = total_ms * 1e3 / (pairs * args.steps)
    roofline.update(stage_common)
    roofline["timing"] = timing_note
    roofline["pipeline_frac"] = sum(floors.values()) / us_per_pair
    roofline["pipeline"] = {"stage_floors_us_per_pair": floors, "measured_us_per_pair": us_per_pair,
                            "what": "sum of the three dominant stages' issue floors at this run's SM clock (graph: 21 "
                                    "fp32 ops per node pair; triangles: algorithmic work at the binding pipe's measured "
                                    "peak; score: 16 FP32 instructions per hypothesis x point) / measured time per pair"}

    # ---- CPU baseline: the from-paper oracle on this box's host cores (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_batched(pairs, ps, cfg, d_i.cpu().numpy())

    if rank == 0:
        score_us_per_step = stage_us["score"] / args.steps
        line = {
            "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{WORKLOAD}: {pairs} pairs/GPU x N={N}, {ratio_label(cfg)} inliers, tau_c={cfg['tau']}"
                            + (" (BASELINE.json configs[1])" if WORKLOAD == "cfg2_3dmatch_256x5000" else ""),
                "pairs_per_gpu": pairs, "pairs_per_step_whole_job": job_pairs, "N": N,
                "K_e": int(reg.params.num_edges), "apex_per_edge": int(reg.params.apex_per_edge),
                "hypotheses_per_pair": K,
                "parallelism": (f"the same {total_pairs}-pair batch, pair b on GPU b mod {world}, no collective" if strong
                                else f"{world} x independent batches, no collective"),
                "triangle_path": "tensor cores (tcgen05 mxf4)" if path_used == 1 else "POPC bitset",
                "node_prune": node_prune,
                "l2": "512 MB flush write between timed steps; per-step workspace (~3 GB) also exceeds the 126 MB L2",
            },
            "hypotheses_per_sec": {
                "whole_pipeline": value * K,
                "score_kernel": pairs * K / (score_us_per_step * 1e-6),
                "what": "whole_pipeline = value x K (every stage included); score_kernel = B x K / time of the "
                        "scoring kernel alone (SURVEY.md 8d), one GPU's share, from the live stage timers",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(src.nbytes + dst.nbytes),
                    "d2h_bytes_per_step": int(pairs * (36 + 12 + 4)), "matches_device_path": same},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "single_pair_latency": latency,
            "recall_vs_ground_truth": recall, "ms_steps": ms_steps, "wall_s_device_region": wall_dev,
            "workspace_bytes": reg.get("workspace_bytes"), "retries": reg.get("retries"),
        }
        emit(line)
    reg.close()


def run_group(args, emit, torch):
    """One process, --gpus N devices, ONE call per step: sac_cot_group_register_packed deals pair b of the batch to
    device b mod N (strong scaling of a fixed batch)."""
    from sac_cot_b200.api import Group
    cfg = synth.CONFIGS[WORKLOAD]
    pairs = args.pairs or cfg["pairs"]
    N = cfg["N"]
    G = args.gpus
    ps, src, dst, offsets = make_batch(range(pairs))
    grp = Group(range(G), tau_compat=cfg["tau"], tau_inlier=cfg["tau"])
    if args.lanes:
        grp.set("lanes", args.lanes)
    if args.chunk_pairs:
        grp.set("chunk_pairs", args.chunk_pairs)
    K = grp.params.num_edges * grp.params.apex_per_edge
    h_src, h_dst = torch.from_numpy(src).pin_memory(), torch.from_numpy(dst).pin_memory()
    h_R = torch.empty((pairs, 3, 3), dtype=torch.float32).pin_memory()
    h_t = torch.empty((pairs, 3), dtype=torch.float32).pin_memory()
    h_i = torch.empty(pairs, dtype=torch.int32).pin_memory()

    def step():
        grp.register_packed_ptr(h_src.data_ptr(), h_dst.data_ptr(), offsets, h_R.data_ptr(), h_t.data_ptr(), h_i.data_ptr())

    sampler = ClockSampler(0)
    for _ in range(args.warmup + 2):
        step()
    launches0 = sum(grp.get(g, "launches") for g in range(G))
    sampler.start()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()   # returns when every device has delivered its results
        ts.append(time.perf_counter() - t0)
    clocks = sampler.stop()
    launches = sum(grp.get(g, "launches") for g in range(G)) - launches0
    value = pairs * args.steps / sum(ts)
    ok = 0
    for b in range(pairs):
        ang, dt_ = synth.pose_error(h_R[b].numpy(), h_t[b].numpy(), ps[b].R_gt, ps[b].t_gt)
        ok += ang < np.deg2rad(5.0) and dt_ < 1.5 * cfg["tau"]
    # the same batch through one device of the group: identical bits, and the 1-device time of the same call
    with Registrar(device=0, tau_compat=cfg["tau"], tau_inlier=cfg["tau"]) as one:
        r1R = torch.empty((pairs, 3, 3), dtype=torch.float32).pin_memory()
        r1t = torch.empty((pairs, 3), dtype=torch.float32).pin_memory()
        r1i = torch.empty(pairs, dtype=torch.int32).pin_memory()
        t1 = []
        for k in range(args.warmup + 2 + args.steps):
            t0 = time.perf_counter()
            one.register_packed_ptr(h_src.data_ptr(), h_dst.data_ptr(), offsets, r1R.data_ptr(), r1t.data_ptr(),
                                    r1i.data_ptr(), _abi.LOC_HOST)
            if k >= args.warmup + 2:
                t1.append(time.perf_counter() - t0)
        same = bool((r1R.numpy() == h_R.numpy()).all() and (r1t.numpy() == h_t.numpy()).all()
                    and (r1i.numpy() == h_i.numpy()).all())
    one_value = pairs * args.steps / sum(t1)
    line = {
        "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(ts) / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: ONE batch of {pairs} pairs x N={N}, {ratio_label(cfg)} inliers, tau_c={cfg['tau']}",
                   "parallelism": f"one process, one sac_cot_group_register_packed call per step: pair b on GPU b mod {G}, "
                                  "one enqueueing thread per device, no collective",
                   "pairs_per_step_whole_job": pairs, "N": N, "hypotheses_per_pair": K,
                   "timing": "wall clock around the call (host buffers in, results out); `value` equals `e2e` because the "
                             "group entry point takes host buffers only"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(src.nbytes + dst.nbytes),
                "d2h_bytes_per_step": int(pairs * 52)},
        "same_call_on_one_gpu": {"value": one_value, "speedup": value / one_value, "bit_identical": same},
        "gpu_launches": int(launches), "clocks": clocks, "recall_vs_ground_truth": ok / pairs,
        "ms_steps": [1e3 * x for x in ts],
    }
    emit(line)
    grp.close()


def run_match_bench(args, emit, torch, dev):
    """--stage match: the correspondence front end on the workload's shape — per pair N source and N
    target keypoints with 33-D FPFH-like descriptors — and the front end + registration chain with a device-resident
    hand-off.  One GPU."""
    cfg = synth.CONFIGS[WORKLOAD]
    pairs = args.pairs or cfg["pairs"]
    N, dim = cfg["N"], 33
    ps = [synth.make_config_pair(WORKLOAD, b) for b in range(pairs)]
    descs = [synth.make_descriptors(p, dim, seed=b) for b, p in enumerate(ps)]
    offs = np.arange(pairs + 1, dtype=np.int64) * N
    lib = load_library()
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    reg = Registrar(lib=lib, device=dev.index, stream=stream.cuda_stream, tau_compat=cfg["tau"], tau_inlier=cfg["tau"])
    T = lambda parts: torch.from_numpy(np.ascontiguousarray(np.concatenate(parts))).to(dev)  # noqa: E731
    d_f, d_g = T([d[0] for d in descs]), T([d[1] for d in descs])
    d_xs, d_xd = T([p.src for p in ps]), T([p.dst for p in ps])
    d_nn = torch.empty(pairs * N, dtype=torch.int32, device=dev)
    d_cs = torch.empty((pairs * N, 3), dtype=torch.float32, device=dev)
    d_cd = torch.empty((pairs * N, 3), dtype=torch.float32, device=dev)
    d_R = torch.empty((pairs, 3, 3), dtype=torch.float32, device=dev)
    d_t = torch.empty((pairs, 3), dtype=torch.float32, device=dev)
    d_i = torch.empty(pairs, dtype=torch.int32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def match():
        reg.match_packed_ptr(d_f.data_ptr(), d_xs.data_ptr(), offs, d_g.data_ptr(), d_xd.data_ptr(), offs, dim,
                             d_nn.data_ptr(), d_cs.data_ptr(), d_cd.data_ptr(), _abi.LOC_DEVICE)

    def register():
        reg.register_packed_ptr(d_cs.data_ptr(), d_cd.data_ptr(), offs, d_R.data_ptr(), d_t.data_ptr(), d_i.data_ptr(),
                                _abi.LOC_DEVICE)

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush.fill_(k & 0xFF)
            ev[k][0].record(stream)
            fn()
            ev[k][1].record(stream)
        torch.cuda.synchronize(dev)
        return [a.elapsed_time(b) for a, b in ev]

    sampler = ClockSampler(dev.index)
    for _ in range(args.warmup):
        match()
        register()
        torch.cuda.synchronize(dev)
        reg.get("last_status")
    sampler.start()
    launches0 = reg.get("launches")
    ms_match = timed(match, args.steps)
    launches = reg.get("launches") - launches0
    ms_chain = timed(lambda: (match(), register()), args.steps)
    clocks = sampler.stop()
    assert reg.get("last_status") == 0
    nn = d_nn.cpu().numpy().reshape(pairs, N)
    inl_found = float(np.mean([(nn[b][ps[b].inlier_idx] == ps[b].inlier_idx).mean() for b in range(pairs)]))
    R_chk, t_chk = d_R.cpu().numpy(), d_t.cpu().numpy()
    ok = sum(1 for b in range(pairs)
             if (lambda e: e[0] < np.deg2rad(5.0) and e[1] < 1.5 * cfg["tau"])(synth.pose_error(R_chk[b], t_chk[b], ps[b].R_gt, ps[b].t_gt)))
    # stage times
    reg.set("stage_timing", 1)
    for _ in range(args.steps):
        match()
    torch.cuda.synchronize(dev)
    stage_us = {s: reg.get(f"stage_us_{s}") / args.steps for s in ("match_prep", "match_sweep", "match_exact")}
    reg.set("stage_timing", 0)
    # the exhaustive CUDA-core scan of the same batch (the library's own alternative path), on a sample
    reg.set("match_path", 0)
    small = min(pairs, 16)
    offs_s = offs[: small + 1]

    def match_small():
        reg.match_packed_ptr(d_f.data_ptr(), d_xs.data_ptr(), offs_s, d_g.data_ptr(), d_xd.data_ptr(), offs_s, dim,
                             d_nn.data_ptr(), d_cs.data_ptr(), d_cd.data_ptr(), _abi.LOC_DEVICE)
    match_small()
    torch.cuda.synchronize(dev)
    ms_scan = timed(match_small, 3)
    nn_scan = d_nn.cpu().numpy()[: small * N].reshape(small, N)
    same = bool((nn_scan == nn[:small]).all())
    reg.set("match_path", 1)
    # end to end: host descriptors in, correspondences out
    h = [torch.from_numpy(np.ascontiguousarray(np.concatenate(x))).pin_memory()
         for x in ([d[0] for d in descs], [p.src for p in ps], [d[1] for d in descs], [p.dst for p in ps])]
    h_nn = torch.empty(pairs * N, dtype=torch.int32).pin_memory()
    h_cs = torch.empty((pairs * N, 3), dtype=torch.float32).pin_memory()
    h_cd = torch.empty((pairs * N, 3), dtype=torch.float32).pin_memory()

    def match_host():
        reg.match_packed_ptr(h[0].data_ptr(), h[1].data_ptr(), offs, h[2].data_ptr(), h[3].data_ptr(), offs, dim,
                             h_nn.data_ptr(), h_cs.data_ptr(), h_cd.data_ptr(), _abi.LOC_HOST)
    match_host()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        match_host()
    e2e = pairs * args.steps / (time.perf_counter() - t0)
    # CPU baseline: the oracle's brute force on a bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        ol = load_oracle(omp=True)
        with Registrar(lib=ol) as o:
            o.set("threads", os.cpu_count() or 1)
            k = min(pairs, 4)
            t0 = time.perf_counter()
            nn_o, _, _, _ = o.match_batch([d[0] for d in descs[:k]], [p.src for p in ps[:k]], [d[1] for d in descs[:k]],
                                          [p.dst for p in ps[:k]])
            dt = time.perf_counter() - t0
            cpu = {"value": k / dt, "unit": "pairs matched/s", "cores": o.get("threads"), "kind": "port",
                   "sample": f"first {k} pairs, from-paper oracle brute force (fp32 chain), OpenMP build",
                   "indices_match_gpu": bool((nn_o.reshape(k, N) == nn[:k]).all())}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    kprime = ((3 * dim + 3 + 15) // 16) * 16
    flop_pair = 2.0 * N * N * kprime                # executed: bf16x3 split operands, K = 3 dim + 3 padded to 16
    sweep_us = stage_us["match_sweep"]
    achieved = flop_pair * pairs / (sweep_us * 1e-6) / 1e12
    bf16 = peaks.get("bf16_tflops_sustained") or 2250.0
    value = pairs * args.steps / (sum(ms_match) * 1e-3)
    line = {
        "metric": f"descriptor matchings/sec at {N} x {N} keypoints, {dim}-D", "value": value, "unit": "pairs matched/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms_match) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3 search, f32 decision",
        "data": "synthetic",
        "config": {"workload": f"front end of {WORKLOAD}: {pairs} pairs x ({N} source, {N} target) keypoints, {dim}-D FPFH-like "
                               "descriptors (synth.make_descriptors)", "stage": "match (SURVEY.md 8f-1)",
                   "l2": "512 MB flush write between timed steps"},
        "e2e": {"value": e2e, "unit": "pairs matched/s", "h2d_bytes_per_step": int(sum(x.numel() * 4 for x in h)),
                "d2h_bytes_per_step": int(pairs * N * 28)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {
            "kernel": "match_mma_kernel (tcgen05 kind::f16, bf16 x bf16 -> f32 in TMEM, M128 N256, candidate epilogue)",
            "bound": "tensor", "achieved": achieved, "peak": bf16, "unit": "TFLOP/s", "frac": achieved / bf16, "traffic": None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md)",
            "note": "K is only 112: a tile is 7 MMAs against 32 768 accumulator values read back and compared, so the "
                    "kernel is bound by the epilogue (tcgen05.ld + 2 instructions per value), not by the tensor pipe; "
                    "`epilogue` relates the compared values to the FP32/ALU lane rate",
            "epilogue": {"values_per_s": float(N) * N * pairs / (sweep_us * 1e-6),
                         "frac_of_lane_rate_at_3_instr": 3.0 * N * N * pairs / (sweep_us * 1e-6) / fp32_lane_rate(clocks.get("sm_mhz") or 1965.0)},
            "stage_us_per_step": stage_us, "executed_flop_per_pair": flop_pair,
        },
        "cpu_baseline": cpu,
        "exhaustive_cuda_core_scan": {"pairs_per_s": small * 3 / (sum(ms_scan) * 1e-3), "sample_pairs": small,
                                      "same_indices": same, "speedup_of_tensor_path": value / (small * 3 / (sum(ms_scan) * 1e-3))},
        "chain_match_then_register": {"registrations_per_s": pairs * args.steps / (sum(ms_chain) * 1e-3),
                                      "ms_per_step": sum(ms_chain) / args.steps, "recall_vs_ground_truth": ok / pairs,
                                      "what": "descriptors -> correspondences -> (R, t), device-resident hand-off"},
        "inliers_matched_to_true_partner": inl_found,
    }
    emit(line)
    reg.close()


def run_single_sharded(args, rank, local_rank, world, emit, torch, dist, dev):
    """cfg5: one N = 50 000 pair through sac_cot_register_sharded (collectives inside the library)."""
    cfg = synth.CONFIGS[WORKLOAD]
    N = cfg["N"]
    p = synth.make_config_pair(WORKLOAD, 0)
    lib = load_library()
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    reg = Registrar(lib=lib, device=local_rank, stream=stream.cuda_stream, tau_compat=cfg["tau"], tau_inlier=cfg["tau"])
    if args.triangle_path >= 0:
        reg.set("triangle_path", args.triangle_path)
    if args.tile_runs >= 0:
        reg.set("tile_runs", args.tile_runs)
    K = reg.params.num_edges * reg.params.apex_per_edge
    if world > 1:
        reg.comm_init()   # rank 0's ncclUniqueId travels over the process group
    else:
        ident = (ctypes.c_ubyte * _abi.COMM_ID_BYTES)()
        assert lib.sac_cot_comm_unique_id(ident) == 0
        reg.comm_init(rank=0, world=1, unique_id=bytes(ident))

    d_src, d_dst = torch.from_numpy(p.src).to(dev), torch.from_numpy(p.dst).to(dev)
    d_R = torch.zeros(9, dtype=torch.float32, device=dev)
    d_t = torch.zeros(3, dtype=torch.float32, device=dev)
    d_i = torch.zeros(1, dtype=torch.int32, device=dev)
    h_src, h_dst = torch.from_numpy(p.src).pin_memory(), torch.from_numpy(p.dst).pin_memory()
    h_R, h_t = torch.empty(9).pin_memory(), torch.empty(3).pin_memory()
    h_i = torch.empty(1, dtype=torch.int32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        reg.register_sharded_ptr(d_src.data_ptr(), d_dst.data_ptr(), N, d_R.data_ptr(), d_t.data_ptr(), d_i.data_ptr(),
                                 _abi.LOC_DEVICE)

    def step_host():
        reg.register_sharded_ptr(h_src.data_ptr(), h_dst.data_ptr(), N, h_R.data_ptr(), h_t.data_ptr(), h_i.data_ptr(),
                                 _abi.LOC_HOST)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup + 1):
        step_device()
        torch.cuda.synchronize(dev)
        reg.get("last_status")
    step_device()
    torch.cuda.synchronize(dev)
    assert reg.get("last_status") == 0
    R_sh, t_sh, i_sh = d_R.cpu().numpy().reshape(3, 3), d_t.cpu().numpy(), int(d_i.item())
    ang, dtr = synth.pose_error(R_sh, t_sh, p.R_gt, p.t_gt)

    launches0 = reg.get("launches")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        if world > 1:
            dist.barrier()      # the ranks enter every registration together, as one collective call would
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = reg.get("launches") - launches0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_max = torch.tensor(ms_steps, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)   # a registration ends when its slowest rank does
    total_ms = float(ms_max.sum().item())
    value = args.steps / (total_ms * 1e-3)

    # per-stage pass
    reg.set("stage_timing", 1)
    for _ in range(args.steps):
        step_device()
    torch.cuda.synchronize(dev)
    stage_us, stage_calls, stage_common = stage_report(reg, args.steps)
    reg.set("stage_timing", 0)
    path_used = reg.get("triangle_path_used")

    # e2e: host buffers in, result out
    step_host()
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = args.steps / float(e2e_s.item())
    same_host = bool((h_R.numpy().reshape(3, 3) == R_sh).all() and (h_t.numpy() == t_sh).all() and int(h_i[0]) == i_sh)

    # parity flag: the unsharded call on this rank's GPU (its own ctx) gives the same bits
    with Registrar(lib=lib, device=local_rank, tau_compat=cfg["tau"], tau_inlier=cfg["tau"]) as one:
        if args.triangle_path >= 0:
            one.set("triangle_path", args.triangle_path)
        R1, t1, i1 = one.register(p.src, p.dst)
        t0 = time.perf_counter()
        one.register(p.src, p.dst)
        unsharded_ms = 1e3 * (time.perf_counter() - t0)
    identical = torch.tensor([1 if ((R1 == R_sh).all() and (t1 == t_sh).all() and i1 == i_sh) else 0], device=dev)
    if world > 1:
        dist.all_reduce(identical, op=dist.ReduceOp.MIN)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    tri_us = stage_us["triangles"] / max(1, stage_calls["triangles"])
    if path_used == 1:
        roofline, _ = tensor_roofline(reg, N, tri_us, 1.0, 1.0 / world, sm_mhz, peaks, None)
        roofline["note"] = f"this rank counts the triangle cells it owns: 1/{world} of the pair's algorithmic flops per launch"
    else:
        roofline = {"kernel": "triangles (POPC bitset)", "bound": "hbm", "achieved": None, "peak": peaks.get("hbm_gbs"),
                    "unit": "GB/s", "frac": None, "traffic": None, "launch_us": tri_us}
    roofline.update(stage_common)
    roofline["timing"] = f"CUDA events around every stage over {args.steps} registrations on rank 0"

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s_all, threads, res = time_oracle(load_oracle(omp=True), [p], cfg, os.cpu_count())
        cpu_baseline = {"value": 1.0 / s_all, "unit": UNIT, "cores": int(threads), "kind": "port",
                        "sample": f"the one N={N} pair, once, from-paper oracle, OpenMP build ({s_all:.1f} s); "
                                  "a single thread would take minutes, so the all-cores figure is the one reported",
                        "inlier_count_matches_gpu": bool(int(res.inliers[0]) == i_sh)}

    if rank == 0:
        line = {
            "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{WORKLOAD}: one pair, N={N}, {ratio_label(cfg)} inliers, tau_c={cfg['tau']} "
                            "(BASELINE.json configs[4])",
                "N": N, "K_e": int(reg.params.num_edges), "apex_per_edge": int(reg.params.apex_per_edge),
                "hypotheses_per_pair": K,
                "parallelism": f"sac_cot_register_sharded over {world} rank(s): triangle cells (1920 cols x 256 rows) and "
                               "hypothesis ranges split; ncclAllGather + ncclAllReduce(max) enqueued by the library",
                "triangle_path": "tensor cores (tcgen05 mxf4)" if path_used == 1 else "POPC bitset",
                "node_prune": node_prune,
                "l2": "512 MB flush write between timed steps; the adjacency alone (313 MB) exceeds the 126 MB L2",
            },
            "ms_per_registration": total_ms / args.steps,
            "bit_identical_to_unsharded_on_every_rank": bool(identical.item()),
            "unsharded_host_call_ms_same_gpu": unsharded_ms,
            "hypotheses_per_sec": {"whole_pipeline": value * K},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(p.src.nbytes + p.dst.nbytes),
                    "d2h_bytes_per_step": 36 + 12 + 4 + 16, "matches_device_path": same_host},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "pose_error_vs_ground_truth": {"rot_deg": float(np.degrees(ang)), "trans": dtr, "inliers": i_sh},
            "ms_steps": [float(x) for x in ms_max.cpu().tolist()],
            "workspace_bytes": reg.get("workspace_bytes"), "retries": reg.get("retries"),
        }
        emit(line)
    reg.close()


def main():
    global WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=None, help="pairs per step (default: the config's; per GPU when weak)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="batched workloads: weak = every GPU its own batch; strong = one batch, pair b on GPU b mod G")
    ap.add_argument("--chunk-pairs", type=int, default=0, help="library knob chunk_pairs (0 = auto)")
    ap.add_argument("--lanes", type=int, default=0, help="library knob lanes (0 = library default)")
    ap.add_argument("--triangle-path", type=int, default=-1,
                    help="library knob triangle_path (0 POPC, 1 tensor core, 2 by edge density = library default)")
    ap.add_argument("--triangle-dbg", type=int, default=0, help="experiments only (library knob triangle_dbg)")
    ap.add_argument("--tile-runs", type=int, default=-1, help="experiments only (library knob tile_runs)")
    ap.add_argument("--node-prune", type=int, default=-1,
                    help="library knob node_prune: 0 = every pair counts all N x N triangles (round-1 behaviour), "
                         "1 = exact node pruning (library default)")
    ap.add_argument("--workload", default=WORKLOAD, choices=sorted(synth.CONFIGS),
                    help="synthetic config (default: the headline config, BASELINE.json configs[1])")
    ap.add_argument("--stage", default="register", choices=["register", "match"],
                    help="register = the hot path (default); match = the descriptor-matching front end on the workload's shape")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    WORKLOAD = args.workload

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line: anything a library prints (e.g. NCCL's version banner)
    # goes to stderr while the benchmark runs
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — sac_cot_b200 has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.stage == "match":
        run_match_bench(args, emit, torch, dev)
    elif world == 1 and args.gpus > 1:
        if torch.cuda.device_count() < args.gpus:
            raise SystemExit(f"bench.py: --gpus {args.gpus} without torchrun needs that many visible devices")
        run_group(args, emit, torch)
    elif synth.CONFIGS[WORKLOAD]["pairs"] == 1 and WORKLOAD != "cfg1_single_n1000":
        run_single_sharded(args, rank, local_rank, world, emit, torch, dist, dev)
    else:
        run_batched(args, rank, local_rank, world, emit, torch, dist, dev)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
