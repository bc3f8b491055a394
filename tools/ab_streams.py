"""dp_streams sweep on one detector (batch 64 and 32 VGA frames, tensor mode): ms per step from CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from partsbaseddetector_b200 import Model, PartsBasedDetector  # noqa: E402
from partsbaseddetector_b200.synth import synth_frames  # noqa: E402

H, W = 480, 640
frames = synth_frames(8, H, W)
det = PartsBasedDetector(device=0, stream=torch.cuda.current_stream().cuda_stream)
det.distributeModel(Model.load_bin(os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm")))
det.set_option("response_mode", 2)
det.set_option("thresh", -1.14)
for B in (64, 32):
    fr = np.ascontiguousarray(np.concatenate([frames] * (B // 8)))
    dev = torch.from_numpy(fr).cuda()
    ref = None
    for ns in (1, 2, 3, 4):
        det.set_option("dp_streams", ns)
        for _ in range(3):
            det.enqueue_device(dev.data_ptr(), B, H, W, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            det.enqueue_device(dev.data_ptr(), B, H, W, 3)
        e1.record()
        torch.cuda.synchronize()
        rv = np.concatenate([det.rootv(f, l).ravel() for f in (0, B - 1) for l in range(det.nscales())])
        if ref is None:
            ref = rv
        print("batch %d dp_streams %d: %.3f ms/step %.1f frames/s  rootv identical to dp_streams=1: %s"
              % (B, ns, e0.elapsed_time(e1) / 10, B * 10 / e0.elapsed_time(e1) * 1e3, bool(np.array_equal(rv, ref))), flush=True)
