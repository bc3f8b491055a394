"""Small end-to-end run through the newer paths for compute-sanitizer: response modes 2 and 3 (tensor kernels), the DP stage's frame
groups on concurrent streams, the device-side sort + NMS and the pipelined submit/collect API.  8 frames of 120x168."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from partsbaseddetector_b200 import Model, PartsBasedDetector  # noqa: E402
from partsbaseddetector_b200.synth import synth_frames  # noqa: E402

frames = synth_frames(8, 120, 168, start=7)
det = PartsBasedDetector(device=0)
det.distributeModel(Model.load_bin(os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm")))
det.set_option("thresh", -1.2)
det.set_option("dp_streams", 2)
counts = []
for mode in (3, 2):
    det.set_option("response_mode", mode)
    for ov in (-1.0, 0.2):
        det.set_option("nms_overlap", ov)
        counts.append(len(det.detect(frames)))
        counts.append(len(det.collect_ticket(det.submit(frames))))
print("candidates per run:", counts)
assert counts[0] == counts[1] and counts[2] == counts[3] and counts[2] <= counts[0]
det.close()
