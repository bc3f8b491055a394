"""Flip statistics of response mode 2 (tensor) against the CPU oracle on the bench's synthetic VGA frames: cells whose root
mixture (rooti) differs, the root-score gap there, candidate identity at a ~60-candidate threshold."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib
from partsbaseddetector_b200 import Model, PartsBasedDetector
from partsbaseddetector_b200.synth import synth_frames

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
G = int(sys.argv[2]) if len(sys.argv) > 2 else 0
path = os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm")
d = PartsBasedDetector(device=0)
d.distributeModel(Model.load_bin(path))
d.set_option("response_mode", int(os.environ.get("PBD_TC_MODE", "3")))      # 2: tf32x3, 3: fp16x3
d.set_option("tc_taps_per_partial", G)
O = oracle_lib.OracleDetector(Model.load_bin(path).to_flat(), 32)
oracle_lib.use_all_cores()
frames = synth_frames(n, 480, 640, start=0)
tot_cells = tot_flip = 0
for i in range(n):
    O.run(frames[i], 1, 3)
    nl = O.nlevels()
    rv = np.sort(np.concatenate([O.rootv(l).ravel() for l in range(nl)]))
    k = rv.size - 60
    thr = float(0.5 * (float(rv[k - 1]) + float(rv[k])))
    O.set_thresh(thr); O.run(None, 4, 4)
    oc = O.candidates()
    d.set_option("thresh", thr)
    cands = d.detect(frames[i])
    same = len(cands) == len(oc) and all(g.level == o["level"] and np.array_equal(g.x, o["x"]) and np.array_equal(g.y, o["y"]) and
                                         np.array_equal(g.m, o["m"]) for g, o in zip(cands, oc))
    flips, worst = 0, 0.0
    for l in range(nl):
        a, b = d.rooti(0, l), O.rooti(l)
        ra, rb = d.rootv(0, l), O.rootv(l)
        worst = max(worst, float(np.abs(ra - rb).max() / np.abs(rb).max()))
        for (y, x) in zip(*np.nonzero(a != b)):
            flips += 1
            print("  frame %d level %d cell (%d,%d): rooti %d vs %d, rootv %.9g vs %.9g (threshold %.6g)" % (i, l, y, x, a[y, x], b[y, x], ra[y, x], rb[y, x], thr))
        tot_cells += a.size
    bp_flips = 0
    for p, m in ((3, 2), (25, 0), (12, 4), (7, 1)):
        gi, oi = d.backptr(0, 0, 0, p, m), O.backptr(0, 0, p, m)
        bp_flips += sum(int((u != v).sum()) for u, v in zip(gi, oi))
    tot_flip += flips
    print("frame %d: candidates %d identical %s, rooti flips %d, back-pointer flips (4 maps, level 0) %d, max rel rootv err %.2e" % (i, len(oc), same, flips, bp_flips, worst))
print("total rooti flips %d of %d cells" % (tot_flip, tot_cells))
