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
# round 2: exact mode with the root-map NMS epilogue, the three dt_pass variants, CUDA-graph replay, the standalone transform through its
# three kernel generations (streaming / parallel-in-q / lagged scan) and the level-0-in-place HOG path on a grey frame
import numpy as np  # noqa: E402
import torch  # noqa: E402
from partsbaseddetector_b200 import Dt2dPlan  # noqa: E402
det.set_option("response_mode", 0)
det.set_option("nms_overlap", -1.0)
base = len(det.detect(frames))
for v in (1, 2, 0):
    det.set_option("dt_variant", v)
    assert len(det.detect(frames)) == base
det.set_option("root_nms", 2)
assert 0 < len(det.detect(frames)) < base
det.set_option("root_nms", 0)
dev = torch.from_numpy(frames).cuda()
det.set_option("graph", 1)
for _ in range(3):
    det.enqueue_device(dev.data_ptr(), 8, 120, 168, 3)
    assert len(det.collect()) == base
det.set_option("graph", 0)
assert len(det.detect(np.ascontiguousarray(frames[:2, :, :, 1]))) >= 0
rng = np.random.default_rng(3)
for (h, w) in ((37, 53), (70, 300)):
    maps = torch.from_numpy(rng.standard_normal((5, h, w)).astype(np.float32)).cuda()
    outs = []
    for impl in (1, 2, 3):
        plan = Dt2dPlan(5, h, w, [0.02, 0.01, 0.015, -0.01], [2, -3], impl)
        o = torch.empty_like(maps)
        ix = torch.empty((5, h, w), dtype=torch.int16, device="cuda")
        iy = torch.empty_like(ix)
        plan.run(maps.data_ptr(), o.data_ptr(), ix.data_ptr(), iy.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs.append((o.cpu(), ix.cpu(), iy.cpu()))
        plan.close()
    assert all(torch.equal(a, b) for other in outs[1:] for a, b in zip(other, outs[0]))
# round 2, second session: the windowed transform (dt_variant 3) with its lines cut into segments -- forced for the 8-frame batch, chosen
# automatically for a single frame -- and the specialised HOG gather (sbin 4)
det.set_option("dt_variant", 3)
ref = [(c.frame, c.level, c.x.tolist(), c.y.tolist()) for c in det.detect(frames)]
for seg in (32, 48, -1):
    det.set_option("dt_segment", seg)
    assert [(c.frame, c.level, c.x.tolist(), c.y.tolist()) for c in det.detect(frames)] == ref
one = [(c.level, c.x.tolist(), c.y.tolist()) for c in det.detect(frames[3])]
assert one == [(l, x, y) for f, l, x, y in ref if f == 3]
print("round-2 paths ok")
det.close()
