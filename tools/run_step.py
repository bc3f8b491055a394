"""Runs a few steps of the hot path (device-resident frames) -- the command profiled under ncu."""
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
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--h", type=int, default=480)
ap.add_argument("--w", type=int, default=640)
ap.add_argument("--fast", action="store_true")
ap.add_argument("--mode", type=int, default=-1, help="response mode 0 exact, 1 ffma, 2 tensor (tf32), 3 tensor (fp16)")
ap.add_argument("--G", type=int, default=0)
ap.add_argument("--max-levels", type=int, default=0)
ap.add_argument("--nms", type=float, default=-1.0, help="device Candidate::sort + NMS with this overlap")
ap.add_argument("--root-nms", type=int, default=0, help="root-map NMS window")
ap.add_argument("--opt", action="append", default=[], help="detector option key=value")
a = ap.parse_args()
frames = synth_frames(min(a.batch, 4), a.h, a.w)
frames = np.ascontiguousarray(np.concatenate([frames] * ((a.batch + 3) // 4))[:a.batch])
dev = torch.from_numpy(frames).cuda()
det = PartsBasedDetector(device=0, stream=torch.cuda.current_stream().cuda_stream)
det.distributeModel(Model.load_bin(os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm")))
det.set_option("exact", 0 if a.fast else 1)
if a.mode >= 0:
    det.set_option("response_mode", a.mode)
    det.set_option("tc_taps_per_partial", a.G)
det.set_option("max_levels", a.max_levels)
det.set_option("thresh", -1.14)
if a.nms >= 0:
    det.set_option("nms_overlap", a.nms)
det.set_option("root_nms", a.root_nms)
for kv in a.opt:
    k, v = kv.split("=")
    det.set_option(k, float(v))
det.set_option("timing", 1)
for _ in range(a.steps):
    det.enqueue_device(dev.data_ptr(), a.batch, a.h, a.w, 3)
torch.cuda.synchronize()
st = det.stage_times_ms()
tot = sum(st.values())
print("batch %d %dx%d levels %d: stage_ms %s total %.3f ms => %.1f frames/s (last step, per-stage events), launches %d"
      % (a.batch, a.h, a.w, det.nscales(), {k: round(v, 3) for k, v in st.items()}, tot, a.batch / tot * 1e3, det.launch_count()))
if any(kv.startswith('dt_variant') for kv in a.opt):
    print('replayed lines (all steps):', det.get_option('dt_replayed_lines'))
det.close()
