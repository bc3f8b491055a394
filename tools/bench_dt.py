"""Config 5 (BASELINE.json): DT microbenchmark, 26 parts x 6 mixtures = 156 fp32 score maps, GB/s sweep over map sizes.
Algorithmic bytes per cell = 4 in + 4 value + 2 Ix + 2 Iy = 12 B (u16 back-pointers; the reference's int32 Mats would be 16 B).
Prints one JSON line per size.  Device-resident inputs, CUDA events, >= 3 warm-ups."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from partsbaseddetector_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=int, nargs="+", default=[256, 512, 1024, 2048, 4096])
ap.add_argument("--maps", type=int, default=156)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
L = _lib.lib()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
rng = np.random.default_rng(4242)
for n in a.sizes:
    nm = a.maps
    defw = np.stack([rng.uniform(0.01, 0.02, nm), rng.uniform(-0.02, 0.02, nm), rng.uniform(0.01, 0.02, nm), rng.uniform(-0.02, 0.02, nm)], 1).astype(np.float32)
    anchors = np.stack([rng.integers(-3, 4, nm), rng.integers(-2, 6, nm)], 1).astype(np.int32)
    g = torch.Generator(device="cuda").manual_seed(4242)
    d_in = torch.randn((nm, n, n), device="cuda", generator=g)
    d_out = torch.empty_like(d_in)
    d_ix = torch.empty((nm, n, n), dtype=torch.int16, device="cuda")
    d_iy = torch.empty((nm, n, n), dtype=torch.int16, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    def run():
        _lib.check(L.pbd_dt2d_f32_device(C.c_void_p(stream), C.c_void_p(d_in.data_ptr()), nm, n, n, np.ascontiguousarray(defw.reshape(-1)),
                                         np.ascontiguousarray(anchors.reshape(-1)), C.c_void_p(d_out.data_ptr()), C.c_void_p(d_ix.data_ptr()),
                                         C.c_void_p(d_iy.data_ptr()), 0))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    byts = 12.0 * nm * n * n
    print(json.dumps({"bench": "dt2d", "maps": nm, "size": n, "ms": ms, "alg_GBps": byts / ms / 1e6, "frac_of_hbm": byts / ms / 1e6 / peaks["hbm_gbs"],
                      "note": "includes cudaMalloc/cudaFree of scratch and the pointer-composition pass inside pbd_dt2d_f32_device"}))
    del d_in, d_out, d_ix, d_iy
    torch.cuda.empty_cache()
