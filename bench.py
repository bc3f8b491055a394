#!/usr/bin/env python
"""bench.py -- the detect() hot path (person model, full HOG pyramid) on N B200s, next to the reference's CPU path.

  python bench.py --gpus N --steps K --warmup W                 # our CUDA path (one process per GPU under torchrun), config `vga`
  python bench.py --impl reference --steps K --warmup W         # the reference's CPU path (restated oracle, all host threads)
  python bench.py --config {vga,vga1,1080p,dt}                  # BASELINE.json configs 3 / 2 / 4 / 5 (one JSON line each)
  python bench.py --total-frames 256 --gpus 8                   # config 3 as stated: 256 frames in total, split over the ranks (strong)

One "step" = one pass of the whole path (image pyramid -> HOG -> part responses -> DT/DP -> backtrack) over one batch of synthetic
BGR frames per GPU.  Frames are independent, so ranks share nothing on the data path; torch.distributed is
used for the barrier and the max-over-ranks of the timing only.  Rank 0 prints ONE JSON line.

The default response arithmetic is `exact` (separately rounded fp32 multiply / add in the reference's order): the only mode whose
candidates equal the CPU reference's on every frame.  `--mode tensor16` is the fast mode (tcgen05 fp16x3 split products, scores
within 1e-6 relative): `parity` counts what differs over every frame of the timed batch, and both modes are always reported
(`other_response_modes`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm")
MODEL_FLAT = os.path.join(ROOT, "tests", "golden", "Person_26parts_flat.npz")
METRIC = "frames/sec (person model, VGA, full pyramid)"
MODES = {"exact": 0, "ffma": 1, "tensor": 2, "tensor16": 3}      # pbd_set_option("response_mode", ...)
MODE_TXT = {"exact": "exact (separately rounded fp32 multiply/add in the reference's order: bit-identical scores and candidates)",
            "ffma": "ffma (FP32 fused multiply-add responses)",
            "tensor": "tensor (tcgen05 tf32x3 split products + fp32 accumulate for the part responses)",
            "tensor16": "tensor16 (tcgen05 kind::f16 MMAs on fp16 hi/lo splits of the pre-scaled fp32 operands, fp32 accumulate)"}
CONFIGS = {   # BASELINE.json `configs`
    "vga": dict(h=480, w=640, batch=64, max_levels=0, graph=0, steps=10, warmup=3,
                workload="config_person.by_parts (Person_26parts), 640x480 BGR frames, full 14-level HOG pyramid, 1xB200 per rank"),
    "vga1": dict(h=480, w=640, batch=1, max_levels=0, graph=1, steps=200, warmup=20,
                 workload="config_person.by_parts (Person_26parts), ONE 640x480 BGR frame per step, full 14-level HOG pyramid, CUDA-graph replay, 1xB200"),
    "1080p": dict(h=1080, w=1920, batch=8, max_levels=10, graph=0, steps=10, warmup=3,
                  workload="config_person.by_parts (Person_26parts), 1920x1080 BGR frames, first 10 pyramid levels, 1xB200"),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for (t, line) in self.lines:
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except ValueError:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------------ CPU arm
def make_oracle(max_levels=0):
    """The restated reference CPU path, built from the committed flat-model fixture: the product library is not involved."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    O = oracle_lib.OracleDetector(oracle_lib.load_flat_npz(MODEL_FLAT), 32)
    if max_levels:
        O.set_max_levels(max_levels)
    return O, oracle_lib.use_all_cores()


def cpu_frames_per_sec(frames, max_levels=0, budget_s=20.0, max_frames=64, warmup=1):
    """Times the restated reference CPU path (oracle, OpenMP over all host threads) on a bounded sample."""
    O, cores = make_oracle(max_levels)
    for i in range(warmup):
        O.run(frames[i % len(frames)])
    t0 = time.time()
    n = 0
    stage = {}
    while n < max_frames and (time.time() - t0 < budget_s or n == 0):
        O.run(frames[n % len(frames)])
        for k, v in O.timings().items():
            stage[k] = stage.get(k, 0.0) + v
        n += 1
    dt = time.time() - t0
    return n / dt, cores, n, dt, {k: 1e3 * v / n for k, v in stage.items()}


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config if args.config in CONFIGS else "vga"]
    from partsbaseddetector_b200.synth import synth_frames
    per_step = 2 if cfg["h"] <= 480 else 1                     # frames per step: a bounded sample of the GPU arm's batch
    frames = synth_frames(per_step, cfg["h"], cfg["w"], start=0)
    O, cores = make_oracle(cfg["max_levels"])
    for _ in range(args.warmup):
        for f in frames:
            O.run(f)
    t0 = time.time()
    for _ in range(args.steps):
        for f in frames:
            O.run(f)
    dt = time.time() - t0
    fps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "frames_per_step": per_step,
                   "impl_detail": "restated reference CPU path (oracle/pbd_oracle.cpp, OpenMP structure of the reference; pinned bit for bit to the reference's own "
                                  "HOGFeatures.cpp / DistanceTransform.hpp / DynamicProgram.cpp compiled in oracle/_ref); the whole reference needs OpenCV C++ / Boost "
                                  "and cannot be built here"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": "%d steps x %d synthetic frames" % (args.steps, per_step)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ parity over the timed batch
def key_of(level, x, y):
    return (int(level), int(x[0]), int(y[0]))


class BatchOracle:
    """The CPU oracle's candidates and root-mixture maps of `frames` at one threshold (computed once, compared with every mode)."""

    def __init__(self, frames, thr, max_levels):
        O, self.cores = make_oracle(max_levels)
        O.set_thresh(thr)
        self.cands, self.rooti, self.cells = [], [], 0
        t0 = time.time()
        for f in frames:
            O.run(f, 1, 4)
            self.cands.append({key_of(o["level"], o["x"], o["y"]): o for o in O.candidates()})
            self.rooti.append([O.rooti(l).copy() for l in range(O.nlevels())])
        self.cells = int(sum(r.size for r in self.rooti[0])) if frames is not None and len(frames) else 0
        self.seconds = time.time() - t0

    def compare(self, det, cl, mode):
        """cl = CandidateList of the CUDA path for the same frames (batch order); det still holds that batch."""
        n = len(self.cands)
        per = [dict() for _ in range(n)]
        for g in cl:
            per[g.frame][key_of(g.level, g.x, g.y)] = g
        S = dict(frames_checked=n, candidates=0, candidates_gpu=0, one_sided=0, integer_outputs_differ=0, max_rel_root_score_error=0.0,
                 scores_bit_identical=True, rooti_cells_differing=0, rooti_cells=self.cells * n)
        for i in range(n):
            oc, gc = self.cands[i], per[i]
            S["candidates"] += len(oc)
            S["candidates_gpu"] += len(gc)
            for k in set(oc) | set(gc):
                if k not in oc or k not in gc:
                    S["one_sided"] += 1
                    continue
                o, g = oc[k], gc[k]
                if not (np.array_equal(g.x, o["x"]) and np.array_equal(g.y, o["y"]) and np.array_equal(g.m, o["m"]) and np.array_equal(g.parts(), o["rects"])):
                    S["integer_outputs_differ"] += 1
                a, b = float(g.score()), float(o["score"])
                S["scores_bit_identical"] &= a == b
                S["max_rel_root_score_error"] = max(S["max_rel_root_score_error"], abs(a - b) / abs(b))
            for l, ro in enumerate(self.rooti[i]):
                S["rooti_cells_differing"] += int((det.rooti(i, l) != ro).sum())
        S["candidates_differing"] = S["one_sided"] + S["integer_outputs_differ"]
        S["checked"] = "every candidate of %d distinct frames of the timed batch vs the CPU oracle: existence, part x / y / mixture id, rects, root score" % n
        S["tolerance"] = "integers identical; root scores within 1e-4 relative (bit-identical in exact mode)"
        bad = S["max_rel_root_score_error"] > 1e-4 or (mode == "exact" and (S["candidates_differing"] or not S["scores_bit_identical"] or S["rooti_cells_differing"]))
        bad = bad or S["candidates_differing"] > max(2, S["candidates"] // 1000)
        if bad:
            raise SystemExit("bench.py parity gate failed in mode %s: %s" % (mode, json.dumps(S)))
        return S


# ------------------------------------------------------------------------------------------------ GPU arm: whole path
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from partsbaseddetector_b200 import Model, PartsBasedDetector
    from partsbaseddetector_b200.synth import synth_frames
    from partsbaseddetector_b200.sharding import frame_range, max_over_ranks

    cfg = CONFIGS[args.config]
    H, W, C = cfg["h"], cfg["w"], 3
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path is the only implementation (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    strong = args.total_frames > 0
    B = args.batch if args.batch > 0 else cfg["batch"]
    if strong:
        if args.total_frames % world:
            raise SystemExit("--total-frames must be a multiple of the number of ranks")
        B = args.total_frames // world
    steps = args.steps if args.steps > 0 else cfg["steps"]
    warmup = max(args.warmup if args.warmup > 0 else cfg["warmup"], 3)
    # every rank gets its own frames (frame-parallel sharding: global frame index = rank*B + i), all distinct by default
    uniq = min(B, args.unique_frames if args.unique_frames > 0 else B)
    base = synth_frames(uniq, H, W, start=frame_range(rank, world, B)[0])
    host = torch.empty((B, H, W, C), dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    for i in range(B):
        hnp[i] = base[i % uniq]
    dev = host.cuda(non_blocking=False)

    det = PartsBasedDetector(device=local, stream=torch.cuda.current_stream().cuda_stream)
    det.distributeModel(Model.load_bin(MODEL))
    det.set_option("max_candidates", 1 << 20)
    det.set_option("max_levels", cfg["max_levels"])
    for kv in args.opt:                                        # experiments: --opt dt_variant=1
        k, v = kv.split("=")
        det.set_option(k, float(v))
    mode = MODES[args.mode]
    det.set_option("response_mode", mode)
    det.set_option("timing", 1)
    # calibrate the detection threshold on the first frame so that ~50 candidates come back from it (synthetic frames score
    # below the model's -0.75: SURVEY.md section 8d); done once, in exact arithmetic, outside every timed region
    det.set_option("response_mode", 0)
    det.set_option("thresh", 1e9)
    det.detect_device(dev.data_ptr(), B, H, W, C)
    nl = det.nscales()
    rv = np.sort(np.concatenate([det.rootv(0, l).ravel() for l in range(nl)]))
    thr = float(0.5 * (float(rv[-51]) + float(rv[-50]))) if args.thresh is None else args.thresh
    det.set_option("thresh", thr)
    det.set_option("response_mode", mode)
    cells = int(sum(det.level_info(l)["oh"] * det.level_info(l)["ow"] for l in range(nl)))

    # ---- parity over the frames of the timed batch (rank 0) ----
    parity, oracle, parity_other = None, None, {}
    if rank == 0 and not args.no_cpu:
        npar = min(uniq, args.parity_frames if args.parity_frames > 0 else (uniq if H <= 480 else 2))
        oracle = BatchOracle(base[:npar], thr, cfg["max_levels"])
        cl = det.detect_device(dev.data_ptr(), B, H, W, C)
        sub = [g for g in cl if g.frame < npar]
        parity = oracle.compare(det, sub, args.mode)
        parity["oracle_seconds"] = round(oracle.seconds, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    det.set_option("graph", cfg["graph"])
    det.set_option("timing", 0 if cfg["graph"] else 1)
    # ---- device-resident throughput: `value` ----
    for _ in range(warmup):
        det.enqueue_device(dev.data_ptr(), B, H, W, C)
    ncand = len(det.collect())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    l0 = det.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(steps):
        det.enqueue_device(dev.data_ptr(), B, H, W, C)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = det.launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    det.set_option("graph", 0)
    # per-stage device time of one more step (events recorded by the library on the same stream)
    det.set_option("timing", 1)
    det.enqueue_device(dev.data_ptr(), B, H, W, C)
    stage_ms = det.stage_times_ms()
    # per-kernel averages (events after every kernel of the pdf / dp_min stages) over a few more steps of the same workload;
    # timing == 2 runs the DP stage on ONE stream (dp_streams is ignored) so that each interval is one kernel alone: the
    # per-kernel sums therefore exceed the dp_min stage time of the timed region, where frame groups overlap
    det.set_option("timing", 2)
    kt = {}
    nk = min(steps, 5)
    for _ in range(nk):
        det.enqueue_device(dev.data_ptr(), B, H, W, C)
        for k, v in det.kernel_times_ms().items():
            kt[k] = kt.get(k, 0.0) + v / nk
    det.set_option("timing", 1)
    ms_max = max_over_ranks(ms, device="cuda")

    # ---- end to end through the public API with host buffers: `e2e` ----
    def e2e_fps(nsteps):
        """public streaming API (PartsBasedDetector.submit / collect_ticket): every step uploads its batch from pinned host memory
        and downloads its candidates; two batches are kept in flight so transfers overlap the compute of the neighbour"""
        det.set_option("timing", 0)          # no per-stage events: lets the library overlap the chunked H2D with pyramid + HOG
        prev = None
        for _ in range(warmup):                  # untimed warm-up of the same pipelined path
            cur = det.submit(hnp)
            if prev is not None:
                det.collect_ticket(prev)
            prev = cur
        det.collect_ticket(prev)
        barrier()
        t0 = time.time()
        nc_total = 0
        prev = None
        for _ in range(nsteps):
            cur = det.submit(hnp)                     # H2D of the batch + all stages, asynchronous
            if prev is not None:
                nc_total += len(det.collect_ticket(prev))   # D2H of hit count and candidates of the previous step
            prev = cur
        nc_total += len(det.collect_ticket(prev))
        torch.cuda.synchronize()
        s = time.time() - t0
        det.set_option("timing", 1)
        return max_over_ranks(s, device="cuda"), nc_total

    e2e_s, nc_total = e2e_fps(steps)
    d2h = 4 + (nc_total // max(steps, 1)) * (24 + 3 * 26 * 4)      # hit count + per hit: Hit record + (x, y, mixture) x 26 parts

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fps = world * B * steps / (ms_max * 1e-3)
        sm_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz", 1965.0)
        line = {
            "metric": METRIC if args.config == "vga" else METRIC.replace("VGA", {"vga1": "one VGA frame, latency", "1080p": "1080p, 10 levels"}[args.config]),
            "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32" if not args.mode.startswith("tensor") else "f32 (part responses as %s split tensor-core products with fp32 accumulation; everything else f32/f64 as the reference)" % ("tf32x3" if args.mode == "tensor" else "fp16x3"),
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "config": args.config,
                       "batch_per_gpu": B, "total_frames_per_step": B * world, "distinct_frames_per_gpu": uniq, "frame": [H, W, C], "levels": nl, "cells_per_frame": cells,
                       "parallelism": "frame-parallel x%d, no collective" % world,
                       "mode": MODE_TXT[args.mode], "thresh": thr, "candidates_per_step": ncand, "dp_streams": int(det.get_option("dp_streams")),
                       "cuda_graph": bool(cfg["graph"]),
                       "l2": "inputs larger than L2 (%.0f MB of responses per step)" % (552.0 * cells * B / 1e6) if B * cells * 552 > 2.6e8 else
                             "working set of one step (%.0f MB) fits the 126 MB L2: latency configuration, not a bandwidth number" % (1500.0 * cells * B / 1e6)},
            "clocks": clocks,
            "e2e": {"value": world * B * steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * C, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "stage_ms": stage_ms, "kernel_ms": kt,
        }
        if parity is not None:
            line["parity"] = parity
        # ---- rooflines.  Algorithmic work per cell from DESIGN.md section 3 / SURVEY.md section 8d ----
        nmaps = 133                                               # (part, mixture) child maps of the person model = DTs per level
        nlaunch_dt = 22                                           # 11 waves x (rows, columns)
        dt_ms = kt.get("dt_rows", 0.0) + kt.get("dt_cols", 0.0)
        dt_bytes = nmaps * 16.0 * cells * B                       # 4 B in + 4 B value + 4 B Ix + 4 B Iy per map cell (both passes together)
        resp_ms = kt.get("part_response", 0.0)
        resp_flops = 2.0 * 800 * 138 * cells * B                  # the reference's multiply-adds
        step_ms = ms_max / steps
        variant = next((kv.split("=")[1] for kv in args.opt if kv.startswith("dt_variant=")), "3")
        win = variant == "3"
        roof_dt = {"kernel": ("dt_pass_win" if win else "dt_pass") + " (22 launches per frame group and step: 11 waves x rows/columns)", "bound": "hbm", "achieved": dt_bytes / (dt_ms * 1e-3) / 1e9 if dt_ms else None,
                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": dt_bytes / (dt_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if dt_ms else None, "traffic": None,
                   "peak_source": peak_src, "ms_per_launch": dt_ms / nlaunch_dt, "share_of_step": dt_ms / step_ms,
                   "algorithmic_bytes": "16 B per map cell (SURVEY 8d) x 133 maps x cells x batch; as implemented 20 B (the row->column intermediate) with u16 pointers",
                   "note": ("windowed certified evaluation (11 candidates per position, certificate per position, replay of uncertifiable lines), one lane "
                            "per line (per line segment for launches that cannot fill the GPU): bounded by instruction issue (74 instructions per position, "
                            "issue slots 68 % busy, ALU pipe 59 %), not by HBM (DESIGN.md section 3.4)") if win else
                           ("sequential lower-envelope scan with fp64 break points, one lane per line: bounded by instruction issue / latency, not by HBM "
                            "(DESIGN.md section 3.2 incl. the measured parallel-in-q alternative)")}
        tp = os.path.join(ROOT, "profiles", "dt_pass_win_traffic.json" if win else "dt_pass_traffic.json")
        if os.path.exists(tp) and H <= 480:
            tr = json.load(open(tp))
            # the capture holds launches of the largest wave: scale to this run's batch and to the average number of maps per launch
            roof_dt["traffic"] = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["launches"] * B / tr["batch"] * (nmaps / 11.0) / tr.get("maps_per_launch", nmaps / 11.0)
            roof_dt["traffic_source"] = tr.get("source", "ncu")
        if args.mode.startswith("tensor"):
            hw_flops = 3.0 * 2 * 800 * 144 * 128 * ((cells * 1.04) // 128) * B      # 3 split products, 144 padded filters, ~4 % strip padding
            f16 = args.mode == "tensor16"
            tc_peak = peaks.get("bf16_tflops", 1590.0) / (1 if f16 else 2)
            roof_resp = {"kernel": "part_response_tc<%s>" % ("f16" if f16 else "tf32"), "bound": "tensor", "achieved": hw_flops / (resp_ms * 1e-3) / 1e12 if resp_ms else None,
                         "peak": tc_peak, "unit": "TFLOP/s", "frac": hw_flops / (resp_ms * 1e-3) / 1e12 / tc_peak if resp_ms else None, "traffic": None,
                         "ms_per_launch": resp_ms,
                         "peak_source": peak_src + (" cuBLAS bf16 burst (kind::f16 issues at the bf16 rate)" if f16 else
                                                    " cuBLAS bf16 burst / 2 (tf32 issues at half the bf16 rate; nominal dense tf32 1100)"),
                         "algorithmic_tflops": resp_flops / (resp_ms * 1e-3) / 1e12 if resp_ms else None,
                         "note": "hardware FLOP/s of the three split products; algorithmic_tflops counts the reference's fp32 multiply-adds once"}
            tp = os.path.join(ROOT, "profiles", "part_response_tc16_traffic.json" if f16 else "part_response_tc_traffic.json")
        else:
            # The contraction is 325 FLOP per algorithmic byte: bounded by the FP32 pipes, not by HBM and (bit-exact rounding) not by the
            # tensor cores -- SURVEY.md section 8(d) asks for exactly this kernel to be reported against the FP32 roofline.
            fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
            ach = resp_flops / (resp_ms * 1e-3) / 1e12 if resp_ms else None
            roof_resp = {"kernel": "part_response<5,5,%s>" % ("exact" if args.mode == "exact" else "ffma"), "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": ach / fp32_peak if ach else None, "traffic": None,
                         "peak_source": "148 SMs x 128 FP32 lanes x 2 FLOP x sm_max_mhz (nominal FFMA rate; a separately rounded FMUL+FADD pair per multiply-add "
                                        "can reach at most 0.46 of it: tools/ubench_fp32.cu, DESIGN.md section 3.1)",
                         "ms_per_launch": resp_ms,
                         "hbm": {"achieved_GBps": 680.0 * cells * B / (resp_ms * 1e-3) / 1e9 if resp_ms else None, "peak_GBps": peaks["hbm_gbs"],
                                 "note": "680 algorithmic bytes per cell: the kernel sits at a few % of the HBM roofline by construction (325 FLOP/B)"}}
            tp = os.path.join(ROOT, "profiles", "part_response_traffic.json")
        if os.path.exists(tp) and H <= 480:
            tr = json.load(open(tp))
            roof_resp["traffic"] = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * B / tr["batch"]
            roof_resp["traffic_source"] = tr.get("source", "ncu")
        roof_resp["share_of_step"] = resp_ms / step_ms
        # `roofline` = the kernel with the largest share of the step; the other one is reported beside it
        if dt_ms >= resp_ms:
            line["roofline"], line["roofline_part_response"] = roof_dt, roof_resp
        else:
            line["roofline"], line["roofline_dt_pass"] = roof_resp, roof_dt
        # the HBM-bound remainder of the step, for context: algorithmic bytes / kernel time
        mm_bytes = (133 * 4.0 + 131 * 9.0) * cells * B
        hog_bytes = (48.0 + 76.0 + 76.0 + 128.0) * cells * B
        pyr_bytes = 2.0 * (stage_ms.get("pyramid", 0) and 1) * 3.0 * 16.8 * cells * B      # ~16.8 pixels per cell, read + written, 3 channels
        line["roofline_other"] = {
            "mix_max (11 launches)": {"alg_GBps": mm_bytes / (kt["mix_max"] * 1e-3) / 1e9 if kt.get("mix_max") else None,
                                      "frac_of_hbm": mm_bytes / (kt["mix_max"] * 1e-3) / 1e9 / peaks["hbm_gbs"] if kt.get("mix_max") else None},
            "hog (hog_hist + hog_feat)": {"alg_GBps": hog_bytes / (stage_ms["hog"] * 1e-3) / 1e9, "frac_of_hbm": hog_bytes / (stage_ms["hog"] * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "pyramid (resize + pyrDown)": {"alg_GBps": pyr_bytes / (stage_ms["pyramid"] * 1e-3) / 1e9 if stage_ms.get("pyramid") else None,
                                           "frac_of_hbm": pyr_bytes / (stage_ms["pyramid"] * 1e-3) / 1e9 / peaks["hbm_gbs"] if stage_ms.get("pyramid") else None},
        }
        # the other response modes on the same workload (rank 0): device-resident, end to end and -- tensor16 -- what differs from the oracle
        others = {}
        if world == 1 or not strong:
            for om in ("exact", "tensor16", "tensor", "ffma"):
                if om == args.mode or (om in ("ffma", "tensor") and not args.also_fast):
                    continue
                det.set_option("response_mode", MODES[om])
                for _ in range(3):
                    det.enqueue_device(dev.data_ptr(), B, H, W, C)
                torch.cuda.synchronize()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(steps):
                    det.enqueue_device(dev.data_ptr(), B, H, W, C)
                f1.record()
                torch.cuda.synchronize()
                fms = f0.elapsed_time(f1)
                others[om] = {"value": B * steps / (fms * 1e-3), "unit": "frames/s", "ms_per_step": fms / steps, "scope": "rank 0, device-resident"}
                if world == 1:
                    s_e2e, _ = e2e_fps(steps)
                    others[om]["e2e"] = {"value": B * steps / s_e2e, "unit": "frames/s"}
                    if oracle is not None:
                        cl = det.detect_device(dev.data_ptr(), B, H, W, C)
                        others[om]["parity"] = oracle.compare(det, [g for g in cl if g.frame < len(oracle.cands)], om)
            det.set_option("response_mode", mode)
        line["other_response_modes"] = others
        if world == 1 and not args.no_cpu:
            cfps, cores, n, dt, cstage = cpu_frames_per_sec(base, cfg["max_levels"], budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "%d synthetic frames of this workload in %.1f s (restated reference CPU path, OpenMP)" % (n, dt), "stage_ms": cstage}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm: DT microbenchmark (config 5)
def run_dt_arm(args):
    """26 parts x 6 mixtures = 156 fp32 score maps, HBM GB/s sweep over map sizes 256^2 ... 4096^2 on one GPU.  Everything is
    pre-allocated (Dt2dPlan): the timed region holds kernel launches only.  GB/s = 16 algorithmic bytes per map cell / time."""
    import torch
    from partsbaseddetector_b200 import Dt2dPlan
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    if env_int("RANK", 0) != 0:
        return 0
    torch.cuda.set_device(env_int("LOCAL_RANK", 0))
    peaks, peak_src = measured_peaks()
    nmaps = args.dt_maps
    rng = np.random.default_rng(4242)
    defw = np.stack([rng.uniform(0.01, 0.02, nmaps), rng.uniform(-0.02, 0.02, nmaps), rng.uniform(0.01, 0.02, nmaps), rng.uniform(-0.02, 0.02, nmaps)], axis=1).astype(np.float32)
    anchors = np.stack([rng.integers(-3, 4, nmaps), rng.integers(-2, 6, nmaps)], axis=1).astype(np.int32)
    steps = args.steps if args.steps > 0 else 5
    warmup = max(args.warmup, 3)
    sweep = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # 256 MB > L2: written between timed iterations of the small sizes
    st = torch.cuda.current_stream().cuda_stream
    IMPLS = ((4, "windowed"), (1, "streaming"), (3, "streaming_scan"), (2, "parallel_in_q"))
    # two input families: "score" = white noise at the scale of the person model's response maps (sigma 0.01 against deformation weights
    # 0.01-0.02: measured on the oracle's responses, whose cell-to-cell differences are as large as their spread) -- what the detector's
    # transform sees; "noise" = unit-variance white noise, 100x rougher (a sample's reach is ~20 cells): the worst case
    for size in [int(s) for s in args.dt_sizes.split(",")]:
        gen = torch.Generator(device="cuda")
        gen.manual_seed(4242 + size)
        d_in = torch.randn((nmaps, size, size), generator=gen, device="cuda", dtype=torch.float32)
        d_out = torch.empty_like(d_in)
        d_ix = torch.empty((nmaps, size, size), dtype=torch.int16, device="cuda")
        d_iy = torch.empty_like(d_ix)
        row = {"size": size, "maps": nmaps, "input_bytes": d_in.numel() * 4}
        for family, scale in (("noise", 1.0), ("score", 0.01)):
            if scale != 1.0:
                d_in.mul_(scale)
            fam = {}
            ref_out = None
            for impl, name in IMPLS:
                if impl == 2 and size > 1024:
                    continue
                plan = Dt2dPlan(nmaps, size, size, defw, anchors, impl)
                for _ in range(warmup):
                    plan.run(d_in.data_ptr(), d_out.data_ptr(), d_ix.data_ptr(), d_iy.data_ptr(), 0, st)
                torch.cuda.synchronize()
                plan.replayed()
                tot = 0.0
                for _ in range(steps):
                    flush.fill_(1)                                            # flush L2 (inputs of the large sizes exceed it anyway)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    plan.run(d_in.data_ptr(), d_out.data_ptr(), d_ix.data_ptr(), d_iy.data_ptr(), 0, st)
                    e1.record()
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
                ms = tot / steps
                gbs = 16.0 * nmaps * size * size / (ms * 1e-3) / 1e9
                fam[name] = {"ms": ms, "GBps": gbs, "frac_of_hbm": gbs / peaks["hbm_gbs"]}
                if impl == 4:
                    fam[name]["lines_replayed_per_run"] = plan.replayed() / steps
                    fam[name]["lines_per_run"] = 2 * nmaps * size
                # the implementations must agree bit for bit (checksum of the value maps and of both arg-max maps)
                chk = (int(d_out.view(torch.int32).to(torch.int64).sum().item()), int(d_ix.to(torch.int64).sum().item()), int(d_iy.to(torch.int64).sum().item()))
                if ref_out is None:
                    ref_out = chk
                fam[name]["equals_first_impl"] = chk == ref_out
                plan.close()
            if family == "noise":
                # size-independent property on a sample: the transform dominates input + penalty at the anchor (exact arithmetic: >=)
                o, i = d_out[0, ::37, ::41].cpu().numpy(), d_in[0].cpu().numpy()
                yy, xx = np.mgrid[0:size:37, 0:size:41]
                ax, ay = int(anchors[0, 0]), int(anchors[0, 1])
                inside = (xx + ax >= 0) & (xx + ax < size) & (yy + ay >= 0) & (yy + ay < size)
                row["property_out_ge_anchor_input"] = bool(np.all(o[inside] >= i[np.clip(yy + ay, 0, size - 1), np.clip(xx + ax, 0, size - 1)][inside] - 1e-6))
            row[family] = fam
        sweep.append(row)
        del d_in, d_out, d_ix, d_iy
        torch.cuda.empty_cache()
    top = sweep[-1]
    names = [n for _, n in IMPLS]
    best = {f: max((top[f].get(k) or {}).get("GBps", 0) for k in names) for f in ("score", "noise")}
    best_name = {f: max(names, key=lambda k: (top[f].get(k) or {}).get("GBps", 0)) for f in ("score", "noise")}
    line = {"metric": "DT microbenchmark: algorithmic HBM GB/s (16 B per map cell), %d score-scale maps of %d^2" % (nmaps, top["size"]), "value": best["score"], "unit": "GB/s",
            "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": top["score"][best_name["score"]]["ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (certified fp32 / fp64 break points)", "data": "synthetic",
            "value_unit_variance_noise": best["noise"], "best_impl": best_name,
            "config": {"workload": "DT microbench: %d fp32 maps per size (26 parts x 6 mixtures), w0,w2~U[0.01,0.02], w1,w3~U[-0.02,0.02], anchors U{-3..3}xU{-2..5}; "
                                   "inputs: N(0, 0.01^2) (the scale of the model's response maps: `value`) and N(0, 1) (worst case: `value_unit_variance_noise`)" % nmaps,
                       "config": "dt", "sizes": [r["size"] for r in sweep], "l2": "256 MB written between timed iterations; inputs of the larger sizes exceed L2"},
            "roofline": {"kernel": "%s (rows + columns + dt2d_compose) at %d^2, score-scale maps" % (best_name["score"], top["size"]), "bound": "hbm", "achieved": best["score"],
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": best["score"] / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                         "unit_variance_noise": {"kernel": best_name["noise"], "achieved": best["noise"], "frac": best["noise"] / peaks["hbm_gbs"]}},
            "all_impls_bit_identical": all(v.get("equals_first_impl", True) for r in sweep for f in ("score", "noise") for v in r[f].values()),
            "sweep": sweep, "gpu_launches": 3 * steps}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: the config's, 10 for vga)")
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="vga", choices=["vga", "vga1", "1080p", "dt"],
                    help="vga: batched VGA frames (the headline, BASELINE configs[1]/[2]); vga1: one VGA frame per step, CUDA-graph replay (latency); "
                         "1080p: 1920x1080, first 10 levels; dt: the 156-map distance-transform sweep")
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (default: the config's, 64 for vga)")
    ap.add_argument("--total-frames", type=int, default=0, help="strong scaling: this many frames per step in total, split evenly over the ranks (config 3: 256)")
    ap.add_argument("--unique-frames", type=int, default=0, help="distinct synthetic frames generated per rank (default: all of the batch)")
    ap.add_argument("--parity-frames", type=int, default=0, help="frames of the timed batch compared with the CPU oracle (default: all at VGA, 2 at 1080p)")
    ap.add_argument("--mode", default="exact", choices=["exact", "tensor16", "tensor", "ffma"],
                    help="part-response arithmetic: exact (default) = bit-identical scores on the FP32 pipes; tensor16 / tensor = tcgen05 fp16x3 / tf32x3 split "
                         "products (scores within 1e-6, what differs is counted in `parity`); ffma = fused multiply-add on the FP32 pipes")
    ap.add_argument("--thresh", type=float, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--also-fast", action="store_true", default=False, help="also time the ffma and tf32 tensor modes (rank 0 only)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--opt", action="append", default=[], help="detector option key=value (pbd_set_option), for experiments")
    ap.add_argument("--dt-maps", type=int, default=156)
    ap.add_argument("--dt-sizes", default="256,512,1024,2048,4096")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps <= 0:
            args.steps = 5
        if args.warmup <= 0:
            args.warmup = 1
        return run_reference_arm(args)
    if args.config == "dt":
        return run_dt_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
