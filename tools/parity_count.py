"""Counts, over N DISTINCT synthetic VGA frames, how many candidates of the CUDA path differ from the CPU oracle, per response mode.

    python tools/parity_count.py --frames 256 --modes tensor16,tensor,exact --out gpurun_out/parity_count.json

One fixed threshold for all frames (the bench's rule: the 50th largest root score of frame 0), so a frame yields ~50-250 candidates.
A candidate is identified by (level, root x, root y); it "differs" when it exists on one side only (a root score on the other side of
the threshold) or when any integer output (part x / y / mixture id / rect) differs.  Test infrastructure: uses the oracle.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

MODES = {"exact": 0, "ffma": 1, "tensor": 2, "tensor16": 3}


def key_of(level, x, y):
    return (int(level), int(x[0]), int(y[0]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--modes", default="tensor16,exact")
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--h", type=int, default=480)
    ap.add_argument("--w", type=int, default=640)
    ap.add_argument("--max-levels", type=int, default=0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import oracle_lib
    from partsbaseddetector_b200 import Model, PartsBasedDetector
    from partsbaseddetector_b200.synth import synth_frames
    path = os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm")
    modes = [m for m in args.modes.split(",") if m]
    det = PartsBasedDetector(device=0)
    det.distributeModel(Model.load_bin(path))
    det.set_option("max_candidates", 400000)
    O = oracle_lib.OracleDetector(Model.load_bin(path).to_flat(), 32)
    if args.max_levels:
        det.set_option("max_levels", args.max_levels)
        O.set_max_levels(args.max_levels)
    cores = oracle_lib.use_all_cores()
    stats = {m: dict(candidates_oracle=0, candidates_gpu=0, one_sided=0, integer_outputs_differ=0, frames_with_difference=0,
                     max_rel_score_err=0.0, rooti_cells_differing=0) for m in modes}
    thr = None
    t0 = time.time()
    cells = 0
    for b0 in range(0, args.frames, args.batch):
        nb = min(args.batch, args.frames - b0)
        frames = synth_frames(nb, args.h, args.w, start=args.start + b0)
        oc_all, rooti_all = [], []
        for i in range(nb):
            O.run(frames[i], 1, 3)
            nl = O.nlevels()
            if thr is None:
                rv = np.sort(np.concatenate([O.rootv(l).ravel() for l in range(nl)]))
                k = rv.size - 50
                thr = float(0.5 * (float(rv[k - 1]) + float(rv[k])))
                det.set_option("thresh", thr)
            O.set_thresh(thr)
            O.run(None, 4, 4)
            oc_all.append({key_of(o["level"], o["x"], o["y"]): o for o in O.candidates()})
            rooti_all.append([O.rooti(l).copy() for l in range(nl)])
            if b0 == 0 and i == 0:
                cells = int(sum(r.size for r in rooti_all[0]))
        for m in modes:
            S = stats[m]
            det.set_option("response_mode", MODES[m])
            cl = det.detect(np.stack(frames))
            per = [dict() for _ in range(nb)]
            for g in cl:
                per[g.frame][key_of(g.level, g.x, g.y)] = g
            for i in range(nb):
                oc, gc = oc_all[i], per[i]
                S["candidates_oracle"] += len(oc)
                S["candidates_gpu"] += len(gc)
                bad = 0
                for k in set(oc) | set(gc):
                    if k not in oc or k not in gc:
                        S["one_sided"] += 1
                        bad += 1
                        continue
                    o, g = oc[k], gc[k]
                    if not (np.array_equal(g.x, o["x"]) and np.array_equal(g.y, o["y"]) and np.array_equal(g.m, o["m"]) and
                            np.array_equal(g.parts(), o["rects"])):
                        S["integer_outputs_differ"] += 1
                        bad += 1
                    S["max_rel_score_err"] = max(S["max_rel_score_err"], abs(float(g.score()) - float(o["score"])) / abs(float(o["score"])))
                S["frames_with_difference"] += bad > 0
                for l, ro in enumerate(rooti_all[i]):
                    S["rooti_cells_differing"] += int((det.rooti(i, l) != ro).sum())
        print("frames %d..%d done (%.0f s): %s" % (b0, b0 + nb - 1, time.time() - t0,
                                                    {m: stats[m]["one_sided"] + stats[m]["integer_outputs_differ"] for m in modes}), flush=True)
    out = {"frames": args.frames, "distinct": True, "frame_shape": [args.h, args.w, 3], "thresh": thr, "oracle_cores": cores, "cells_per_frame": cells,
           "max_levels": args.max_levels, "modes": {}}
    for m in modes:
        S = stats[m]
        S["candidates_differing"] = S["one_sided"] + S["integer_outputs_differ"]
        S["rooti_cells"] = cells * args.frames
        out["modes"][m] = S
    print(json.dumps(out))
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
