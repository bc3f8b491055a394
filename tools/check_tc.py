"""Quick device check of response mode 2 (tensor cores) against mode 0 (exact): response error, candidate identity, timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from partsbaseddetector_b200 import Model, PartsBasedDetector
from partsbaseddetector_b200.synth import synth_frame

name = sys.argv[1] if len(sys.argv) > 1 else "Person_26parts"
shape = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (144, 200)
m = Model.load_bin(os.path.join(ROOT, "tests", "golden", name + ".pbdm"))
d = PartsBasedDetector(device=0)
d.distributeModel(m)
img = synth_frame(11, *shape)
nf = m.to_flat().nfilters()
d.set_option("response_mode", 0)
d.pyramid(img); d.pdf()
nl = d.nscales()
ref = [[d.response(0, l, f).copy() for f in range(nf)] for l in range(nl)]
d.set_option("response_mode", 2)
d.set_option("tc_taps_per_partial", int(os.environ.get("TC_G", "0")))
d.pyramid(img); d.pdf()
worst = 0.0
for l in range(nl):
    for f in range(nf):
        a, b = d.response(0, l, f), ref[l][f]
        if not np.all(np.isfinite(a)):
            print("non-finite at level", l, "filter", f); sys.exit(1)
        e = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
        if e > worst:
            worst = e; wl = (l, f)
print("%s %s: levels %d filters %d, worst max-abs error relative to map max: %.3e at %s" % (name, shape, nl, nf, worst, wl))
a, b = d.response(0, 0, 0), ref[0][0]
print("level 0 filter 0 corner tensor:", a[:2, :4].ravel(), "exact:", b[:2, :4].ravel())
pos, neg = [], []
for l in range(min(nl, 3)):
    for f in range(nf):
        a, b = d.response(0, l, f).ravel().astype(np.float64), ref[l][f].ravel().astype(np.float64)
        ulp = np.spacing(np.abs(ref[l][f].ravel()).astype(np.float32)).astype(np.float64)
        e = (a - b) / ulp
        big = np.abs(b) > 0.02
        pos.append(e[(b > 0) & big]); neg.append(e[(b < 0) & big])
pos, neg = np.concatenate(pos), np.concatenate(neg)
print("error in ulps: positive responses mean %.2f std %.2f (n=%d); negative responses mean %.2f std %.2f (n=%d)" %
      (pos.mean(), pos.std(), pos.size, neg.mean(), neg.std(), neg.size))
