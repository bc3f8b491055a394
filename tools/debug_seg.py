"""Debug helper: where do the segmented and the unsegmented windowed walk differ?  (GPU)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from partsbaseddetector_b200 import Model, PartsBasedDetector
from partsbaseddetector_b200.synth import synth_frame
import oracle_lib
name = sys.argv[1] if len(sys.argv) > 1 else "Face_frontal_sparse"
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (120, 160)
fm = Model.load_bin(os.path.join(ROOT, "tests", "golden", name + ".pbdm")).to_flat()
img = synth_frame(21, h, w)
res = {}
for seg in (0, 32, 48):
    d = PartsBasedDetector(device=0)
    d.distributeModel(Model.load_bin(os.path.join(ROOT, "tests", "golden", name + ".pbdm")))
    d.set_option("dt_segment", seg)
    d.get_option("dt_replayed_lines")
    d.detect(img)
    out = {}
    for l in range(d.nscales()):
        for c in range(len(fm.comps)):
            out[("rootv", l, c)] = d.rootv(0, l, c)
            for p in range(1, len(fm.comps[c])):
                for pm in range(fm.nmix(c, fm.comps[c][p].parentid)):
                    ix, iy, ik = d.backptr(0, l, c, p, pm)
                    out[("ix", l, c, p, pm)] = ix; out[("iy", l, c, p, pm)] = iy; out[("ik", l, c, p, pm)] = ik
    res[seg] = out
    print("seg", seg, "replayed", d.get_option("dt_replayed_lines"), "levels", [(d.level_info(l)) for l in range(min(4, d.nscales()))])
    d.close()
for seg in (32, 48):
    nbad = 0
    for k in res[0]:
        a, b = res[0][k], res[seg][k]
        if not np.array_equal(a, b):
            nbad += 1
            if nbad <= 12:
                yy, xx = np.nonzero(a != b)
                print("seg", seg, k, "shape", a.shape, "ndiff", len(yy), "first", list(zip(yy[:6].tolist(), xx[:6].tolist())), a[yy[0], xx[0]], b[yy[0], xx[0]])
    print("seg", seg, "differing maps", nbad, "of", len(res[0]))
