"""Lock-step (warp) execution model of the 1-D envelope variants on REAL score maps (CPU only): the host build of dt_envelope.cuh counts,
per line and sample step, the iterations of the pop loop, the emissions and the cursor advances; a warp of 32 consecutive lines executes
the maximum over its lanes of every loop.  Prints the warp-level iteration counts per step and an instruction estimate from per-iteration
costs read off the SASS (eager: main 78, pop 62, emission 8 + 16 per iteration -- reproduces the 155 instructions per step that ncu
measured; scan: main 86, pop 66, emission 13, advance 12).  Usage: python tools/dt_warp_model.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
import test_dt_envelope_host as T  # noqa: E402
from partsbaseddetector_b200 import Model  # noqa: E402
from partsbaseddetector_b200.synth import synth_frame  # noqa: E402

L = T.envlib()
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
L.envh_stats.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, i32p, i32p, i32p]

fm = Model.load_bin(os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm")).to_flat()
O = oracle_lib.OracleDetector(fm, 32)
O.run(synth_frame(1234, 480, 640), 1, 2)            # pyramid + HOG + responses of one synthetic VGA frame
tot = {0: np.zeros(6), 2: np.zeros(6)}
nsteps = 0
for level in (0, 1, 3, 6):
    for part in (25, 20, 13, 7):                     # leaf-ish parts: their child scores are plain responses
        P = fm.comps[0][part]
        for mm, (fid, did) in enumerate(zip(P.filterid, P.defid)):
            resp = np.ascontiguousarray(O.response(level, fid))
            w = fm.defs[did]
            ax = int(fm.anchors[did][0])
            nl, N = resp.shape
            for variant in (0, 2):
                pops, emits, advs = (np.zeros((nl, N), np.int32) for _ in range(3))
                L.envh_stats(resp, nl, N, float(w[0]), float(w[1]), ax, variant, pops, emits, advs)
                nw = nl // 32
                if nw == 0:
                    continue
                for a, k in ((pops, 0), (emits, 1), (advs, 2)):
                    blk = a[: nw * 32].reshape(nw, 32, N)
                    tot[variant][k] += blk.max(axis=1).sum()          # warp-level iterations
                    tot[variant][3 + k] += blk.mean(axis=1).sum()     # lane-average iterations
                if variant == 0:
                    nsteps += nw * N
for variant, name in ((0, "eager (envelope_stream)"), (2, "lagged scan (envelope_scan<4>)")):
    t = tot[variant] / nsteps
    print("%-32s per warp step: pop iterations %.2f (lane average %.3f), emissions %.2f (%.3f), cursor advances %.2f (%.3f)"
          % (name, t[0], t[3], t[1], t[4], t[2], t[5]))
    if variant == 0:
        est = 78 + 62 * t[0] + 8 + 16 * t[1]
    else:
        est = 86 + 66 * t[0] + 13 * t[1] + 12 * t[2]
    print("%-32s estimated warp instructions per step: %.0f" % ("", est))

# ---- parallel-in-q schedule: how much sequential work is left? ----
i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C")
L.envh_dt1d_parallel.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, f32p, u16p, i64p]
samples = sites = pops = alive = cand = 0
warp_sites = warp_pops = warps = 0
for level in (0, 1, 3, 6):
    for part in (25, 20, 13, 7):
        P = fm.comps[0][part]
        for fid, did in zip(P.filterid, P.defid):
            resp = np.ascontiguousarray(O.response(level, fid))
            w = fm.defs[did]
            nl, N = resp.shape
            per = np.zeros((nl, 4), np.int64)
            dst, ptr = np.empty(N, np.float32), np.empty(N, np.uint16)
            for i in range(nl):
                L.envh_dt1d_parallel(resp[i], 1, N, float(w[0]), float(w[1]), int(fm.anchors[did][0]), dst, ptr, per[i])
            samples += nl * N; sites += per[:, 0].sum(); pops += per[:, 1].sum(); alive += per[:, 2].sum(); cand += per[:, 3].sum()
            nw = nl // 32
            if nw:
                blk = per[: nw * 32].reshape(nw, 32, 4)
                warp_sites += blk[:, :, 0].max(axis=1).sum(); warp_pops += blk[:, :, 1].max(axis=1).sum(); warps += nw
                wsamples = nw * N if level == 0 and part == 25 and fid == P.filterid[0] else None
print("parallel-in-q schedule: pop sites %.3f per sample (candidates s_q <= s_{q-1}: %.3f), pop iterations %.3f per sample, surviving entries %.3f per sample"
      % (sites / samples, cand / samples, pops / samples, alive / samples))
print("lane per line, lock step over the i-th site of 32 lines: %.1f site rounds and %.1f pop iterations per warp and line "
      "(lane average %.1f sites per line)" % (warp_sites / warps, warp_pops / warps, sites / (samples / 158.0) if samples else 0))
