#!/usr/bin/env python
"""bench.py -- frames/sec of the detect() hot path (person model, VGA, full pyramid) on N B200s.

  python bench.py --gpus N --steps K --warmup W          # our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (restated oracle, all host threads)

One "step" = one pass of the whole path (image pyramid -> HOG -> part responses -> DT/DP -> backtrack) over one
batch of synthetic 640x480 BGR frames per GPU.  Weak scaling: every rank processes its own batch; frames are
independent, so there is no collective on the data path.  Rank 0 prints ONE JSON line.
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
H, W, C = 480, 640, 3
METRIC = "frames/sec (person model, VGA, full pyramid)"
MODES = {"exact": 0, "ffma": 1, "tensor": 2, "tensor16": 3}      # pbd_set_option("response_mode", ...)


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
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_frames_per_sec(frames, budget_s=20.0, max_frames=8, warmup=1):
    """Times the restated reference CPU path (oracle, OpenMP over all host threads) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from partsbaseddetector_b200 import Model
    fm = Model.load_bin(MODEL).to_flat()
    O = oracle_lib.OracleDetector(fm, 32)
    cores = oracle_lib.use_all_cores()
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
    from partsbaseddetector_b200.synth import synth_frames
    per_step = 2                                   # frames per step: a bounded sample of the 64-frame workload
    frames = synth_frames(per_step, H, W, start=0)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from partsbaseddetector_b200 import Model
    O = oracle_lib.OracleDetector(Model.load_bin(MODEL).to_flat(), 32)
    cores = oracle_lib.use_all_cores()
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
        "config": {"workload": "config_person.by_parts (Person_26parts), 640x480 BGR frames, full 14-level HOG pyramid", "frames_per_step": per_step,
                   "impl_detail": "restated reference CPU path (oracle/pbd_oracle.cpp, OpenMP as the reference); the reference itself needs OpenCV C++/Boost and cannot be built here"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": "%d steps x %d synthetic VGA frames" % (args.steps, per_step)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def parity_gate(det, frame, mode):
    """One synthetic VGA frame through the CUDA path and the CPU oracle: integer outputs (part locations, mixture ids, rects, root
    mixture maps) must be identical, root scores within 1e-4 relative (north star); bit-identical in exact mode.  Fails loudly."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from partsbaseddetector_b200 import Model
    O = oracle_lib.OracleDetector(Model.load_bin(MODEL).to_flat(), 32)
    oracle_lib.use_all_cores()
    O.run(frame, 1, 3)
    nl = O.nlevels()
    rv = np.sort(np.concatenate([O.rootv(l).ravel() for l in range(nl)]))
    k = rv.size - 60
    thr = float(0.5 * (float(rv[k - 1]) + float(rv[k])))          # between two neighbouring root scores: ~60 candidates
    O.set_thresh(thr)
    O.run(None, 4, 4)
    oc = O.candidates()
    det.set_option("thresh", thr)
    cands = det.detect(frame)
    worst, flips, ncell = 0.0, 0, 0
    for l in range(nl):
        ref, got = O.rootv(l), det.rootv(0, l)
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
        flips += int((det.rooti(0, l) != O.rooti(l)).sum())      # root-mixture arg-max map: can only differ at score near-ties
        ncell += ref.size
    same = len(cands) == len(oc) and all(g.level == o["level"] and np.array_equal(g.x, o["x"]) and np.array_equal(g.y, o["y"]) and
                                         np.array_equal(g.m, o["m"]) and np.array_equal(g.parts(), o["rects"]) for g, o in zip(cands, oc))
    if not same or worst > 1e-4 or (mode == "exact" and (worst != 0.0 or flips)):
        raise SystemExit("bench.py parity gate failed: identical candidates %s, max relative root-score error %.3g, root-mixture flips %d" % (same, worst, flips))
    return {"checked": "1 synthetic VGA frame, all 14 levels, vs the CPU oracle", "candidates": len(oc), "candidate_integer_outputs_identical": True,
            "max_rel_root_score_error": worst, "tolerance": 1e-4, "root_mixture_map_cells_differing": flips, "root_mixture_map_cells": ncell}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from partsbaseddetector_b200 import Model, PartsBasedDetector
    from partsbaseddetector_b200.synth import synth_frames

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path is the only implementation (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    # every rank gets its own frames (frame-parallel sharding: global frame index = rank*B + i)
    uniq = min(B, args.unique_frames)
    from partsbaseddetector_b200.sharding import frame_range
    base = synth_frames(uniq, H, W, start=frame_range(rank, world, B)[0])
    host = torch.empty((B, H, W, C), dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    for i in range(B):
        hnp[i] = base[i % uniq]
    dev = host.cuda(non_blocking=False)

    det = PartsBasedDetector(device=local, stream=torch.cuda.current_stream().cuda_stream)
    det.distributeModel(Model.load_bin(MODEL))
    mode = MODES[args.mode]
    det.set_option("response_mode", mode)
    det.set_option("timing", 1)
    # calibrate the detection threshold on the first batch so that ~50 candidates/frame come back (synthetic
    # frames score below the model's -0.75: SURVEY.md section 8d); done once, outside every timed region
    det.set_option("thresh", 1e9)
    det.detect_device(dev.data_ptr(), B, H, W, C)
    nl = det.nscales()
    rv = np.concatenate([det.rootv(0, l).ravel() for l in range(nl)])
    thr = float(np.sort(rv)[-50]) if args.thresh is None else args.thresh
    det.set_option("thresh", thr)
    cells = int(sum(det.level_info(l)["oh"] * det.level_info(l)["ow"] for l in range(nl)))
    parity = parity_gate(det, base[0], args.mode) if rank == 0 and not args.no_cpu else None
    det.set_option("thresh", thr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: `value` ----
    for _ in range(args.warmup):
        det.enqueue_device(dev.data_ptr(), B, H, W, C)
    ncand = len(det.collect())
    stage_acc = {}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    l0 = det.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        det.enqueue_device(dev.data_ptr(), B, H, W, C)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = det.launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    # per-stage device time of the last step (events recorded by the library on the same stream)
    stage_ms = det.stage_times_ms()
    # per-kernel averages (events after every kernel of the pdf / dp_min stages) over a few more steps of the same workload;
    # timing == 2 runs the DP stage on ONE stream (dp_streams is ignored) so that each interval is one kernel alone: the
    # per-kernel sums therefore exceed the dp_min stage time of the timed region, where frame groups overlap
    det.set_option("timing", 2)
    kt = {}
    nk = min(args.steps, 5)
    for _ in range(nk):
        det.enqueue_device(dev.data_ptr(), B, H, W, C)
        for k, v in det.kernel_times_ms().items():
            kt[k] = kt.get(k, 0.0) + v / nk
    det.set_option("timing", 1)
    from partsbaseddetector_b200.sharding import max_over_ranks
    ms_max = max_over_ranks(ms, device="cuda")

    # ---- end to end through the public API with host buffers: `e2e` ----
    det.set_option("timing", 0)          # no per-stage events: lets the library overlap the chunked H2D with pyramid + HOG
    # public streaming API (PartsBasedDetector.submit / collect_ticket): every step uploads its batch from pinned host
    # memory and downloads its candidates; two batches are kept in flight so transfers overlap the compute of the neighbour
    prev = None
    for _ in range(args.warmup):                  # untimed warm-up of the same pipelined path
        cur = det.submit(hnp)
        if prev is not None:
            det.collect_ticket(prev)
        prev = cur
    det.collect_ticket(prev)
    barrier()
    t0 = time.time()
    nc_total = 0
    prev = None
    for _ in range(args.steps):
        cur = det.submit(hnp)                     # H2D of the batch + all stages, asynchronous
        if prev is not None:
            nc_total += len(det.collect_ticket(prev))   # D2H of hit count and candidates of the previous step
        prev = cur
    nc_total += len(det.collect_ticket(prev))
    torch.cuda.synchronize()
    e2e_s = time.time() - t0
    e2e_s = max_over_ranks(e2e_s, device="cuda")
    d2h = 4 + (nc_total // max(args.steps, 1)) * (24 + 3 * 26 * 4)      # hit count + per hit: Hit record + (x, y, mixture) x 26 parts

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fps = world * B * args.steps / (ms_max * 1e-3)
        sm_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz", 1965.0)
        mode_txt = {"exact": "exact (separately rounded multiply/add, bit-identical scores)", "ffma": "ffma (FP32 fused multiply-add responses)",
                    "tensor": "tensor (tcgen05 tf32x3 split products + fp32 accumulate for the part responses; scores within 2e-6 relative, "
                              "integer outputs identical to the CPU oracle -- checked in `parity`)",
                    "tensor16": "tensor16 (tcgen05 kind::f16 MMAs on fp16 hi/lo splits of the power-of-two pre-scaled fp32 operands: the same 11+11 "
                                "significand bits and the same three products as tf32x3 at half the MMAs and operand bytes, fp32 accumulate; scores "
                                "within 2e-6 relative, integer outputs identical to the CPU oracle -- checked in `parity`)"}[args.mode]
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if not args.mode.startswith("tensor") else "f32 (part responses as %s split tensor-core products with fp32 accumulation; everything else f32/f64 as the reference)" % ("tf32x3" if args.mode == "tensor" else "fp16x3"),
            "data": "synthetic",
            "config": {"workload": "config_person.by_parts (Person_26parts), 640x480 BGR frames, full 14-level HOG pyramid, 1xB200 per rank",
                       "batch_per_gpu": B, "frame": [H, W, C], "levels": nl, "cells_per_frame": cells, "parallelism": "frame-parallel x%d, no collective" % world,
                       "mode": mode_txt, "thresh": thr, "candidates_per_step": ncand, "dp_streams": int(det.get_option("dp_streams")),
                       "l2": "inputs larger than L2 (%.0f MB of responses per step)" % (552.0 * cells * B / 1e6)},
            "clocks": clocks,
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * C, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "stage_ms": stage_ms, "kernel_ms": kt,
        }
        if parity is not None:
            line["parity"] = parity
        # ---- rooflines.  Algorithmic work per cell from DESIGN.md section 3 / SURVEY.md section 8d ----
        nmaps = 133                                               # (part, mixture) child maps of the person model = DTs per level
        nlaunch_dt = 22                                           # 11 waves x (rows, columns)
        dt_ms = kt.get("dt_rows", 0.0) + kt.get("dt_cols", 0.0)
        dt_bytes = 2 * nmaps * 10.0 * cells * B                   # per pass and map cell: 4 B read, 4 B value + 2 B arg-max written
        resp_ms = kt.get("part_response", 0.0)
        resp_flops = 2.0 * 800 * 138 * cells * B                  # the reference's multiply-adds
        roof_dt = {"kernel": "dt_pass (22 launches per step: 11 waves x rows/columns)", "bound": "hbm", "achieved": dt_bytes / (dt_ms * 1e-3) / 1e9 if dt_ms else None,
                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": dt_bytes / (dt_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if dt_ms else None, "traffic": None,
                   "peak_source": peak_src, "ms_per_launch": dt_ms / nlaunch_dt, "share_of_step": dt_ms / (ms_max / args.steps),
                   "note": "sequential lower-envelope scan with fp64 break points, one lane per line: bounded by instruction issue (63 % of issue slots busy, "
                           "62 % lane utilisation in the ncu capture), not by HBM"}
        tp = os.path.join(ROOT, "profiles", "dt_pass_traffic.json")
        if os.path.exists(tp):
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
            fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
            ach = 680.0 * cells * B / (resp_ms * 1e-3) / 1e9 if resp_ms else None
            roof_resp = {"kernel": "part_response", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"] if ach else None,
                         "traffic": None, "peak_source": peak_src, "ms_per_launch": resp_ms,
                         "note": "dense contraction (325 FLOP/B): FP32-issue-bound, not HBM-bound; see fp32",
                         "fp32": {"achieved_tflops": resp_flops / (resp_ms * 1e-3) / 1e12 if resp_ms else None, "peak_tflops": fp32_peak,
                                  "frac": resp_flops / (resp_ms * 1e-3) / 1e12 / fp32_peak if resp_ms else None, "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz"}}
            tp = os.path.join(ROOT, "profiles", "part_response_traffic.json")
        if os.path.exists(tp):
            tr = json.load(open(tp))
            roof_resp["traffic"] = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * B / tr["batch"]
            roof_resp["traffic_source"] = tr.get("source", "ncu")
        roof_resp["share_of_step"] = resp_ms / (ms_max / args.steps)
        # `roofline` = the kernel with the largest share of the step; the other one is reported beside it
        if dt_ms >= resp_ms:
            line["roofline"], line["roofline_part_response"] = roof_dt, roof_resp
        else:
            line["roofline"], line["roofline_dt_pass"] = roof_resp, roof_dt
        # the HBM-bound remainder of the step, for context: algorithmic bytes / kernel time
        mm_bytes = (133 * 4.0 + 131 * 9.0) * cells * B
        hog_bytes = (48.0 + 76.0 + 76.0 + 128.0) * cells * B
        line["roofline_other"] = {
            "mix_max (11 launches)": {"alg_GBps": mm_bytes / (kt["mix_max"] * 1e-3) / 1e9 if kt.get("mix_max") else None,
                                      "frac_of_hbm": mm_bytes / (kt["mix_max"] * 1e-3) / 1e9 / peaks["hbm_gbs"] if kt.get("mix_max") else None},
            "hog (hog_hist + hog_feat)": {"alg_GBps": hog_bytes / (stage_ms["hog"] * 1e-3) / 1e9, "frac_of_hbm": hog_bytes / (stage_ms["hog"] * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        }
        # the other response modes on the same workload (single rank, device-resident), for comparison
        others = {}
        for om in ("exact", "ffma", "tensor", "tensor16"):
            if om == args.mode or (om == "ffma" and not args.also_fast):
                continue
            det.set_option("response_mode", MODES[om])
            for _ in range(3):
                det.enqueue_device(dev.data_ptr(), B, H, W, C)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(args.steps):
                det.enqueue_device(dev.data_ptr(), B, H, W, C)
            f1.record()
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1)
            others[om] = {"value": B * args.steps / (fms * 1e-3), "unit": "frames/s", "ms_per_step": fms / args.steps}
        det.set_option("response_mode", mode)
        line["other_response_modes"] = others
        if world == 1 and not args.no_cpu:
            cfps, cores, n, dt, cstage = cpu_frames_per_sec(base, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "%d synthetic VGA frames in %.1f s (restated reference CPU path, OpenMP)" % (n, dt), "stage_ms": cstage}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--unique-frames", type=int, default=8, help="distinct synthetic frames generated per rank (tiled to the batch)")
    ap.add_argument("--mode", default="tensor16", choices=["tensor16", "tensor", "exact", "ffma"],
                    help="part-response arithmetic: tensor16 = tcgen05 fp16x3 (default) / tensor = tcgen05 tf32x3 (both: scores within 2e-6, integer "
                         "outputs identical to the oracle), exact = bit-identical scores on the FP32 pipes, ffma = fused multiply-add on the FP32 pipes")
    ap.add_argument("--thresh", type=float, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--also-fast", action="store_true", default=False, help="also time the fused-multiply-add mode (rank 0 only)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
