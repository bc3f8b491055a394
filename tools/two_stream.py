"""Experiment: the same frames as ONE batch on one stream vs split over S detectors on S streams (kernel tails of one
sub-batch overlap the other's work).  Prints frames/s for each arrangement (CUDA events, device-resident frames)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from partsbaseddetector_b200 import Model, PartsBasedDetector  # noqa: E402
from partsbaseddetector_b200.synth import synth_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--mode", type=int, default=3)
a = ap.parse_args()
H, W = 480, 640
frames = synth_frames(8, H, W)
frames = np.ascontiguousarray(np.concatenate([frames] * (a.batch // 8))[:a.batch])
dev = torch.from_numpy(frames).cuda()
model = Model.load_bin(os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm"))


def run(nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    dets = []
    for s in streams:
        d = PartsBasedDetector(device=0, stream=s.cuda_stream)
        d.distributeModel(model)
        d.set_option("response_mode", a.mode)
        d.set_option("thresh", -1.14)
        dets.append(d)
    per = a.batch // nstreams
    def step():
        for i, d in enumerate(dets):
            d.enqueue_device(dev.data_ptr() + i * per * H * W * 3, per, H, W, 3)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(a.steps):
        step()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("streams %d x %d frames: %.3f ms/step  %.1f frames/s" % (nstreams, per, ms / a.steps, a.batch * a.steps / ms * 1e3), flush=True)
    for d in dets:
        d.close()


for n in (1, 2, 4):
    run(n)
