"""Wall-clock breakdown of the pipelined API (submit / collect_ticket) vs the synchronous detect()."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from partsbaseddetector_b200 import Model, PartsBasedDetector
from partsbaseddetector_b200.synth import synth_frames
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = synth_frames(8, 480, 640)
host = torch.empty((B, 480, 640, 3), dtype=torch.uint8, pin_memory=True)
hnp = host.numpy()
for i in range(B):
    hnp[i] = base[i % 8]
det = PartsBasedDetector(device=0, stream=torch.cuda.current_stream().cuda_stream)
det.distributeModel(Model.load_bin(os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm")))
det.set_option("thresh", -1.143)
for _ in range(3):
    det.detect(hnp)
torch.cuda.synchronize()
t = time.time()
for _ in range(6):
    c = det.detect(hnp)
torch.cuda.synchronize()
print("sync detect: %.2f ms/step, %d candidates" % ((time.time() - t) / 6 * 1e3, len(c)))
ts, tc = [], []
prev = None
t = time.time()
for i in range(8):
    a = time.time(); cur = det.submit(hnp); b = time.time(); ts.append(b - a)
    if prev is not None:
        a = time.time(); n = len(det.collect_ticket(prev)); tc.append(time.time() - a)
    prev = cur
det.collect_ticket(prev)
torch.cuda.synchronize()
print("pipelined: %.2f ms/step; submit ms %s; collect ms %s" % ((time.time() - t) / 8 * 1e3, [round(x * 1e3, 2) for x in ts], [round(x * 1e3, 2) for x in tc]))
det.close()
