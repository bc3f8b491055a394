"""Host-side logic that needs no GPU: sharding across ranks (gloo, world_size 2), the C++ adapter header, synthetic inputs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from partsbaseddetector_b200.sharding import frame_range, split_frames
from partsbaseddetector_b200.synth import synth_frame, synth_frames


def test_frame_ranges_are_disjoint_and_cover():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += list(frame_range(r, world, 32))
        assert seen == list(range(32 * world))
        parts = split_frames(256 + 3, world)
        assert sum(len(p) for p in parts) == 259 and max(map(len, parts)) - min(map(len, parts)) <= 1
        assert [i for p in parts for i in p] == list(range(259))
    with pytest.raises(ValueError):
        frame_range(2, 2, 8)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from partsbaseddetector_b200.sharding import frame_range, max_over_ranks, sum_over_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = frame_range(rank, world, 4)
    # every rank generates only its own frames; the "timing" of rank r is r+1 -> max over ranks must be `world`
    digest = int(sum(int(synth_frame(i, 24, 32).sum()) for i in frames))
    tmax = max_over_ranks(rank + 1.0)
    total = sum_over_ranks(len(frames))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, list(frames), digest, tmax, total))


def test_two_rank_sharding_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2, 3] and res[1][1] == [4, 5, 6, 7]
    assert res[0][3] == res[1][3] == 2.0 and res[0][4] == res[1][4] == 8.0
    ref = [int(sum(int(synth_frame(i, 24, 32).sum()) for i in r)) for r in ([0, 1, 2, 3], [4, 5, 6, 7])]
    assert [res[0][2], res[1][2]] == ref


def test_synthetic_frames_are_reproducible():
    a, b = synth_frame(5, 48, 64), synth_frames(2, 48, 64, start=4)[1]
    assert a.dtype == np.uint8 and a.shape == (48, 64, 3) and np.array_equal(a, b)
    assert a.std() > 10                      # non-degenerate gradients


def test_cpp_adapter_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = str(tmp_path / "demo")
    libdir = os.path.join(ROOT, "partsbaseddetector_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "demo.cpp"),
                           "-L" + libdir, "-lpbd_b200", "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 255 and "Usage" in r.stdout                       # -1, src/demo.cpp:58-61
    r = subprocess.run([exe, "model.txt", "x.ppm"], capture_output=True, text=True)
    assert r.returncode == 254 and "Unsupported model format" in r.stdout    # -2
    r = subprocess.run([exe, str(tmp_path / "missing.mat"), "x.ppm"], capture_output=True, text=True)
    assert r.returncode == 253 and "Error deserializing" in r.stdout         # .mat goes to MatlabIOModel (src/demo.cpp:69-70)
    r = subprocess.run([exe, str(tmp_path / "missing.xml"), "x.ppm"], capture_output=True, text=True)
    assert r.returncode == 253 and "Error deserializing" in r.stdout         # -3
    import torch
    if not torch.cuda.is_available():
        ppm = tmp_path / "f.ppm"
        im = synth_frame(3, 120, 160)[:, :, ::-1]
        ppm.write_bytes(b"P6\n160 120\n255\n" + np.ascontiguousarray(im).tobytes())
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm"), str(ppm)], capture_output=True, text=True)
        assert r.returncode == 251 and "no CUDA device" in r.stdout          # the CUDA path is the only implementation


def test_stage_plugins_compile_against_the_reference_interfaces(tmp_path):
    """include/pbd_b200_plugins.hpp: CudaHOGFeatures : IFeatures and CudaConvolutionEngine : IConvolutionEngine compile (`override`
    on every virtual) against the interface mirror AND, where /root/reference exists, against the reference's own
    include/IFeatures.hpp / include/IConvolutionEngine.hpp; without a GPU the plug-ins fail loudly."""
    libdir = os.path.join(ROOT, "partsbaseddetector_b200")
    base = ["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "ref_shim"),
            os.path.join(ROOT, "examples", "plugin_demo.cpp"), "-L" + libdir, "-lpbd_b200", "-Wl,-rpath," + libdir]
    exe = str(tmp_path / "plugin_demo")
    subprocess.check_call(base + ["-Wall", "-Werror", "-o", exe])
    if os.path.isdir("/root/reference/include"):
        subprocess.check_call(base + ["-w", "-I/root/reference/include", "-o", str(tmp_path / "plugin_demo_ref")])
    import torch
    if not torch.cuda.is_available():
        raw = tmp_path / "f.raw"
        raw.write_bytes(synth_frame(3, 60, 80).tobytes())
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "Person_26parts.pbdm"), str(raw), "60", "80", "3", str(tmp_path / "o.bin")],
                           capture_output=True, text=True)
        assert r.returncode == 251 and "no CUDA device" in r.stdout


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                                  text=True, timeout=600)
    import json
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "frames/s"


def test_nms_matches_oracle_restatement():
    """Candidate::nonMaximaSuppression (include/Candidate.hpp:277-304): library implementation vs the oracle's restatement
    on random candidate sets (host code: no GPU needed)."""
    import oracle_lib
    from partsbaseddetector_b200 import Candidate
    from partsbaseddetector_b200.detector import CandidateList
    rng = np.random.default_rng(5)
    L = oracle_lib.lib()
    for trial in range(6):
        n, nparts, h, w = int(rng.integers(1, 120)), int(rng.integers(1, 9)), 120, 160
        rects = np.zeros((n, nparts, 4), np.int32)
        cx, cy = rng.integers(-20, w + 20, n), rng.integers(-20, h + 20, n)
        for p in range(nparts):
            rects[:, p, 0] = cx + rng.integers(-15, 15, n)
            rects[:, p, 1] = cy + rng.integers(-15, 15, n)
            rects[:, p, 2] = rng.integers(5, 30, n)
            rects[:, p, 3] = rng.integers(5, 30, n)
        scores = np.sort(rng.standard_normal(n).astype(np.float32))[::-1].copy()
        overlap = float([0.0, 0.1, 0.5][trial % 3])
        keep = np.zeros(n, np.int32)
        nk = L.orc_nms(np.ascontiguousarray(rects).reshape(-1), n, nparts, h, w, overlap, keep)
        meta = np.zeros((n, 4), np.int32)
        meta[:, 3] = nparts
        parts = np.zeros((n, nparts, 7), np.int32)
        parts[:, :, 3:7] = rects
        parts[:, :, 0] = np.arange(n)[:, None]            # tag candidates by index in x[0]
        out = Candidate.nonMaximaSuppression((h, w), CandidateList(meta, scores, parts), overlap)
        assert [int(c.x[0]) for c in out] == keep[:nk].tolist()
        lst = list(CandidateList(meta, scores, parts))
        Candidate.nonMaximaSuppression(np.zeros((h, w, 3), np.uint8), lst, overlap)
        assert [int(c.x[0]) for c in lst] == keep[:nk].tolist()
    # boundingBox
    c = CandidateList(meta, scores, parts)[0]
    x0, y0 = rects[0, :, 0].min(), rects[0, :, 1].min()
    assert c.boundingBox() == (x0, y0, (rects[0, :, 0] + rects[0, :, 2]).max() - x0, (rects[0, :, 1] + rects[0, :, 3]).max() - y0)


def test_batch_sort_then_nms_keeps_frames_independent():
    """The documented batch flow Candidate.sort (a global score sort that interleaves frames) followed by nonMaximaSuppression must
    equal per-frame sort + NMS: the scratch image is per frame, whatever the list order."""
    from partsbaseddetector_b200 import Candidate
    from partsbaseddetector_b200.detector import CandidateList
    rng = np.random.default_rng(11)
    h, w, nparts, nframes = 90, 120, 3, 4
    n = 200
    meta = np.zeros((n, 4), np.int32)
    meta[:, 0] = rng.integers(0, nframes, n)
    meta[:, 3] = nparts
    parts = np.zeros((n, nparts, 7), np.int32)
    parts[:, :, 3] = rng.integers(-10, w, (n, nparts)); parts[:, :, 4] = rng.integers(-10, h, (n, nparts))
    parts[:, :, 5] = rng.integers(4, 25, (n, nparts)); parts[:, :, 6] = rng.integers(4, 25, (n, nparts))
    parts[:, :, 0] = np.arange(n)[:, None]
    scores = rng.permutation(n).astype(np.float32)               # distinct
    # the advisor's minimal case: 2 frames x 2 identical boxes, interleaved by the sort, overlap 0 -> one survivor per frame
    m2 = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 0, 0, 1], [1, 0, 0, 1]], np.int32)
    p2 = np.zeros((4, 1, 7), np.int32); p2[:, 0, 3:7] = (10, 10, 20, 20)
    out = Candidate.nonMaximaSuppression((h, w), CandidateList(m2, np.array([4, 3, 2, 1], np.float32), p2), 0.0)
    assert [(c.frame, float(c.score())) for c in out] == [(0, 4.0), (1, 3.0)]
    for overlap in (0.0, 0.3):
        batch = Candidate.nonMaximaSuppression((h, w), CandidateList(*(a[np.argsort(-scores, kind="stable")] for a in (meta, scores, parts))), overlap)
        got = sorted(int(c.x[0]) for c in batch)
        want = []
        for f in range(nframes):
            sel = np.nonzero(meta[:, 0] == f)[0]
            sel = sel[np.argsort(-scores[sel], kind="stable")]
            want += [int(c.x[0]) for c in Candidate.nonMaximaSuppression((h, w), CandidateList(meta[sel], scores[sel], parts[sel]), overlap)]
        assert got == sorted(want) and 0 < len(got) < n
        assert [float(c.score()) for c in batch] == sorted((float(c.score()) for c in batch), reverse=True)   # list order survives


def test_depth_pruning_matches_oracle_restatement():
    """pbd_candidates_filter_by_depth (SearchSpacePruning::filterCandidatesByDepth, src/SearchSpacePruning.cpp:73-95) vs the oracle's
    restatement -- itself pinned to the reference's compiled source in test_oracle_ref.py -- on random part boxes of the person tree,
    boxes crossing the image border included (both clip)."""
    import oracle_lib
    from conftest import golden_model_path, load_flat
    from partsbaseddetector_b200 import Model, filterCandidatesByDepth
    from partsbaseddetector_b200.detector import CandidateList
    fm = load_flat("Person_26parts")
    model = Model.load_bin(golden_model_path("Person_26parts"))
    comp = fm.comps[0]
    parent = np.array([p.parentid for p in comp], np.int32)
    anchor0 = np.array([fm.anchors[p.defid[0]] if p.defid else (0, 0) for p in comp], np.int32).reshape(-1)
    rng = np.random.default_rng(21)
    h, w, n, nparts = 120, 160, 150, len(comp)
    yy, xx = np.mgrid[0:h, 0:w]
    depth = (2.0 + 0.002 * xx + 0.5 * (yy > 70) + 0.003 * rng.standard_normal((h, w))).astype(np.float32)
    depth[rng.random((h, w)) < 0.2] = 0.0
    rects = np.zeros((n, nparts, 4), np.int32)
    cx, cy = rng.integers(-10, w, n), rng.integers(-10, h, n)
    for p in range(nparts):
        rects[:, p, 0] = cx + rng.integers(-12, 12, n); rects[:, p, 1] = cy + rng.integers(-12, 12, n)
        rects[:, p, 2] = rng.integers(4, 24, n); rects[:, p, 3] = rng.integers(4, 24, n)
    meta = np.zeros((n, 4), np.int32); meta[:, 3] = nparts
    parts = np.zeros((n, nparts, 7), np.int32); parts[:, :, 3:7] = rects; parts[:, :, 0] = np.arange(n)[:, None]
    scores = rng.standard_normal(n).astype(np.float32)
    seen = set()
    for zf in (0.002, 0.03, 0.5):
        keep = np.zeros(n, np.int32)
        k = oracle_lib.lib().orc_filter_by_depth(np.ascontiguousarray(rects).reshape(-1), n, nparts, parent, anchor0, depth.reshape(-1), h, w, zf, keep)
        out = filterCandidatesByDepth(model, CandidateList(meta, scores, parts), depth, zf)
        assert [int(c.x[0]) for c in out] == np.nonzero(keep)[0].tolist() and len(out) == k
        seen.add(k)
    assert len(seen) > 1 and max(seen) > 0                       # the factor matters


def test_pyramid_geometry_matches_oracle_over_many_sizes():
    """Level tables (image sizes, HOG cell counts, scales) of the product's host code vs the oracle's restatement of
    src/HOGFeatures.cpp:95-127,174-176 for ~1500 image sizes incl. exact powers of two of 5*sbin (floor(log/log) edge cases)."""
    import oracle_lib
    from partsbaseddetector_b200 import _lib
    L, O = _lib.lib(), oracle_lib.lib()
    rng = np.random.default_rng(1)
    sizes = [(480, 640), (1080, 1920), (240, 320), (20, 20), (40, 40), (80, 80), (160, 160), (320, 320), (640, 640), (1280, 1280), (2560, 2560),
             (40, 999), (159, 161), (161, 159), (21, 4000)]
    sizes += [(int(h), int(w)) for h, w in zip(rng.integers(20, 2200, 700), rng.integers(20, 2200, 700))]
    for sbin, interval in ((4, 3), (8, 3), (4, 10), (8, 10)):
        for (h, w) in sizes:
            if min(h, w) < 5 * sbin:
                continue
            dims = np.zeros(4 * 96, np.int32)
            sc = np.zeros(96, np.float32)
            n = L.pbd_pyramid_geometry(h, w, sbin, interval, 0, 96, dims, sc)
            wh = np.zeros(2 * 96, np.int32)
            osc = np.zeros(96, np.float32)
            on = O.orc_pyramid_geometry(h, w, sbin, interval, 96, wh, osc)
            assert n == on, (h, w, sbin, interval)
            for l in range(min(n, 96)):
                assert (dims[4 * l + 1], dims[4 * l]) == (wh[2 * l], wh[2 * l + 1]), (h, w, l)
                assert sc[l] == osc[l]
                oh, ow = oracle_lib.C.c_int(), oracle_lib.C.c_int()
                O.orc_hog_dims(int(dims[4 * l]), int(dims[4 * l + 1]), sbin, oracle_lib.C.byref(oh), oracle_lib.C.byref(ow))
                assert (dims[4 * l + 2], dims[4 * l + 3]) == (oh.value, ow.value)
    assert L.pbd_pyramid_geometry(0, 10, 4, 3, 0, 0, dims, sc) < 0
