"""Parity of the CUDA path (through the C-ABI) with the CPU oracle on the same seeded inputs.
Bit-exact for every integer output (pyramid pixels, back-pointers, mixture ids, part locations, rects) and,
in the default exact mode, for every float score as well; the fast mode is held to 1e-5 relative."""
import os

import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN, golden_model_path, load_flat
from partsbaseddetector_b200 import Model, PartsBasedDetector, PbdError, dt2d
from partsbaseddetector_b200.synth import synth_frame, synth_frames, synth_score_map

pytestmark = pytest.mark.gpu

_det_cache = {}


def detector(name):
    if name not in _det_cache:
        d = PartsBasedDetector(device=0)
        d.distributeModel(Model.load_bin(os.path.join(GOLDEN, name + ".pbdm")))
        _det_cache[name] = d
    d = _det_cache[name]
    for k, v in (("exact", 1), ("backptr", 0), ("max_levels", 0), ("thresh", load_flat(name).thresh), ("dt_variant", 3), ("dt_segment", -1), ("root_nms", 0), ("graph", 0)):
        d.set_option(k, v)
    return d


def oracle(name, precision=32):
    return oracle_lib.OracleDetector(load_flat(name), precision)


def lowered_threshold(O, keep=60):
    rv = np.concatenate([O.rootv(l, c).ravel() for l in range(O.nlevels()) for c in range(len(O.model.comps))])
    return float(np.sort(rv)[-min(keep, rv.size)])


def cand_key(c):
    return (c["frame"], c["level"], c["component"], int(c["y"][0]), int(c["x"][0]))


# ------------------------------------------------------------------------------------------ pyramid + HOG
@pytest.mark.parametrize("shape", [(120, 160), (240, 320), (97, 131), (203, 77)])
def test_pyramid_and_hog_bitexact(shape):
    img = synth_frame(shape[0] + shape[1], *shape)
    d, O = detector("Person_26parts"), oracle("Person_26parts")
    d.pyramid(img)
    O.run(img, 1, 1)
    assert d.nscales() == O.nlevels()
    for l in range(O.nlevels()):
        a, b = d.level_info(l), O.level_info(l)
        assert a == b
        assert np.array_equal(d.pyramid_image(0, l), O.image(l)), "level %d image" % l
        assert np.array_equal(d.features(0, l), O.features(l)), "level %d features" % l


def test_hog_gray_and_sbin8():
    img = synth_frame(5, 192, 256)[:, :, 1].copy()
    d, O = detector("Person_8parts"), oracle("Person_8parts")          # sbin 8
    d.pyramid(img)
    O.run(img, 1, 1)
    for l in range(O.nlevels()):
        assert np.array_equal(d.pyramid_image(0, l, channels=1), O.image(l))
        assert np.array_equal(d.features(0, l), O.features(l))


def test_degenerate_frames():
    d, O = detector("Person_26parts"), oracle("Person_26parts")
    for img in (np.zeros((96, 128, 3), np.uint8), np.full((96, 128, 3), 255, np.uint8),
                np.tile(np.arange(128, dtype=np.uint8)[None, :, None] * 2, (96, 1, 3))):
        d.pyramid(img)
        O.run(img, 1, 1)
        for l in range(O.nlevels()):
            assert np.array_equal(d.features(0, l), O.features(l))


# ------------------------------------------------------------------------------------------ responses
@pytest.mark.parametrize("name", ["Person_26parts", "Willowcoffee_5parts", "Face_frontal_sparse", "Person_8parts"])
def test_responses_exact_and_fast(name):
    fm = load_flat(name)
    img = synth_frame(11, 144, 200)
    d, O = detector(name), oracle(name)
    O.run(img, 1, 2)
    d.pyramid(img)
    d.pdf()
    nl = O.nlevels()
    for l in (0, nl // 2, nl - 1):
        for f in sorted(set([0, 1, fm.nfilters() // 2, fm.nfilters() - 1])):
            assert np.array_equal(d.response(0, l, f), O.response(l, f)), (l, f)
    d.set_option("exact", 0)
    d.pyramid(img)
    d.pdf()
    for l in (0, nl - 1):
        for f in (0, fm.nfilters() - 1):
            ref = O.response(l, f)
            assert np.allclose(d.response(0, l, f), ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())


def test_responses_injected_random_features_all_filters():
    # stage-isolated: random features (not HOG-like) on ragged level sizes, every filter compared
    name = "Willowcoffee_5parts"
    fm = load_flat(name)
    d, O = detector(name), oracle(name)
    ohow = [[9, 35], [33, 8], [4, 4], [17, 65]]
    scales = [4.0, 5.0, 6.0, 8.0]
    d.set_levels(1, ohow, scales)
    O.set_levels(ohow, scales)
    rng = np.random.default_rng(0)
    for l, (oh, ow) in enumerate(ohow):
        f = rng.standard_normal((oh, ow, 32)).astype(np.float32)
        d.set_features(0, l, f)
        O.set_features(l, f)
    O.run(None, 2, 2)
    d.pdf()
    for l in range(len(ohow)):
        for f in range(fm.nfilters()):
            assert np.array_equal(d.response(0, l, f), O.response(l, f)), (l, f)


# ------------------------------------------------------------------------------------------ distance transform
@pytest.mark.parametrize("h,w", [(1, 1), (1, 17), (23, 1), (37, 53), (160, 158), (200, 330)])
@pytest.mark.parametrize("mode", [0, 1])
def test_dt2d_bitexact(h, w, mode):
    rng = np.random.default_rng(h * 1000 + w)
    n = 5
    maps = np.stack([synth_score_map(i, h, w) for i in range(n)])
    maps[3] = 0.25                                                    # flat map: exact ties everywhere
    maps[4] = np.round(maps[4] * 2) / 2                               # heavily quantised: many ties
    defw = np.stack([rng.uniform(0.01, 0.02, n), rng.uniform(-0.02, 0.02, n), rng.uniform(0.01, 0.02, n),
                     rng.uniform(-0.02, 0.02, n)], axis=1).astype(np.float32)
    anchors = np.stack([rng.integers(-3, 4, n), rng.integers(-2, 6, n)], axis=1).astype(np.int32)
    out, ix, iy = dt2d(maps, defw, anchors, mode)
    L = oracle_lib.lib()
    for i in range(n):
        o, x, y = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
        L.orc_dt2d_f32(maps[i].reshape(-1), h, w, defw[i], int(anchors[i, 0]), int(anchors[i, 1]), mode, o.reshape(-1), x.reshape(-1), y.reshape(-1))
        assert np.array_equal(out[i], o), i
        assert np.array_equal(ix[i], x), i
        assert np.array_equal(iy[i], y), i


@pytest.mark.parametrize("impl", [1, 2, 3, 4])
@pytest.mark.parametrize("h,w", [(1, 1), (3, 2), (33, 65), (118, 158), (268, 478), (60, 700)])
def test_dt2d_both_kernel_generations_bitexact(h, w, impl):
    """The streaming envelope (impl 1: one lane per line) and the parallel-in-q kernels (impl 2: a warp per batch of lines in shared
    memory; straight-line variants up to 256 samples, loops beyond), the streaming envelope with lagged-scan emission (impl 3) and the
    windowed certified evaluation (impl 4; maps whose anchors exceed the window replay every line) against the oracle, through the
    pre-allocated plan API."""
    import torch
    from partsbaseddetector_b200 import Dt2dPlan
    rng = np.random.default_rng(h * 977 + w + impl)
    n = 7
    maps = np.stack([synth_score_map(10 + i, h, w) for i in range(n)])
    maps[4] = -1.5
    maps[5] = np.round(maps[5] * 2) / 2
    maps[6] = np.cumsum(np.cumsum(maps[6], axis=0), axis=1) * 0.01            # smooth ramps: long runs without pops, then bursts
    defw = np.stack([rng.uniform(0.01, 0.08, n), rng.uniform(-0.02, 0.02, n), rng.uniform(0.01, 0.08, n), rng.uniform(-0.02, 0.02, n)], axis=1).astype(np.float32)
    anchors = np.stack([rng.integers(-8, 9, n), rng.integers(-11, 13, n)], axis=1).astype(np.int32)
    plan = Dt2dPlan(n, h, w, defw, anchors, impl)
    assert plan.impl() == impl
    d_in = torch.from_numpy(maps).cuda()
    d_out = torch.empty_like(d_in)
    d_ix = torch.empty((n, h, w), dtype=torch.int16, device="cuda")
    d_iy = torch.empty_like(d_ix)
    L = oracle_lib.lib()
    for mode in (0, 1):
        plan.run(d_in.data_ptr(), d_out.data_ptr(), d_ix.data_ptr(), d_iy.data_ptr(), mode, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out, ix, iy = d_out.cpu().numpy(), d_ix.cpu().numpy().view(np.uint16), d_iy.cpu().numpy().view(np.uint16)
        for i in range(n):
            o, x, y = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
            L.orc_dt2d_f32(maps[i].reshape(-1), h, w, defw[i], int(anchors[i, 0]), int(anchors[i, 1]), mode, o.reshape(-1), x.reshape(-1), y.reshape(-1))
            assert np.array_equal(out[i], o), (i, mode)
            assert np.array_equal(ix[i].astype(np.int32), x), (i, mode)
            assert np.array_equal(iy[i].astype(np.int32), y), (i, mode)
    plan.close()


@pytest.mark.parametrize("h,w", [(118, 158), (300, 1100), (2050, 64)])
def test_dt2d_windowed_on_score_like_maps(h, w):
    """impl 4 where it is meant to run: maps at the scale of the person model's responses (sigma 0.01 against deformation weights
    0.01-0.02, anchors within the window): bit-identical to the oracle with (almost) no line replayed; plus one map of unit-variance
    noise in the same plan, which is replayed wholesale and still exact."""
    import torch
    from partsbaseddetector_b200 import Dt2dPlan
    rng = np.random.default_rng(h + 31 * w)
    n = 6
    maps = (rng.standard_normal((n, h, w)) * 0.01).astype(np.float32)
    maps[1] += (np.add.outer(np.sin(np.arange(h) / 9.0), np.cos(np.arange(w) / 13.0)) * 0.2).astype(np.float32)
    maps[5] = rng.standard_normal((h, w)).astype(np.float32)
    defw = np.stack([rng.uniform(0.01, 0.02, n), rng.uniform(-0.02, 0.02, n), rng.uniform(0.01, 0.02, n), rng.uniform(-0.02, 0.02, n)], axis=1).astype(np.float32)
    anchors = np.stack([rng.integers(-3, 4, n), rng.integers(-2, 6, n)], axis=1).astype(np.int32)
    plan = Dt2dPlan(n, h, w, defw, anchors, 4)
    d_in = torch.from_numpy(maps).cuda()
    d_out = torch.empty_like(d_in)
    d_ix = torch.empty((n, h, w), dtype=torch.int16, device="cuda")
    d_iy = torch.empty_like(d_ix)
    plan.replayed()
    plan.run(d_in.data_ptr(), d_out.data_ptr(), d_ix.data_ptr(), d_iy.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    replayed = plan.replayed()
    out, ix, iy = d_out.cpu().numpy(), d_ix.cpu().numpy().view(np.uint16), d_iy.cpu().numpy().view(np.uint16)
    L = oracle_lib.lib()
    for i in range(n):
        o, x, y = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
        L.orc_dt2d_f32(maps[i].reshape(-1), h, w, defw[i], int(anchors[i, 0]), int(anchors[i, 1]), 0, o.reshape(-1), x.reshape(-1), y.reshape(-1))
        assert np.array_equal(out[i], o), i
        assert np.array_equal(ix[i].astype(np.int32), x) and np.array_equal(iy[i].astype(np.int32), y), i
    assert (h + w) * 0.5 <= replayed <= (h + w) * 1.6, replayed       # the noise map's lines (most of them), hardly any of the others
    plan.close()


@pytest.mark.parametrize("seg", [16, 32, 48])
@pytest.mark.parametrize("h,w", [(37, 53), (118, 158), (64, 300)])
def test_dt2d_segmented_walk_on_adversarial_maps(h, w, seg):
    """The segmented walk (lines cut into segments of seg - 10 positions, each walked by its own lane; a line is accepted only if all
    of its segments are) on the inputs built to break a windowed certificate: score-like maps, maps quantised to a coarse binary grid
    with weights 2^-k so that break points fall exactly on integers and neighbouring candidates tie exactly, constant maps, a map with
    isolated huge spikes (owners far outside the window), noise, a NaN and infinities of either sign, anchors up to the window's limit.  Every map
    must equal the oracle bit for bit, exactly as with one lane per line."""
    import torch
    from partsbaseddetector_b200 import Dt2dPlan
    rng = np.random.default_rng(h * 131 + w + seg)
    n = 10
    maps = (rng.standard_normal((n, h, w)) * 0.01).astype(np.float32)
    maps[1] += (np.add.outer(np.sin(np.arange(h) / 7.0), np.cos(np.arange(w) / 11.0)) * 0.2).astype(np.float32)
    maps[2] = np.round(rng.standard_normal((h, w)) * 2) / 4                    # quantised: exact ties, integer break points (weights 2^-k below)
    maps[3] = np.round(rng.standard_normal((h, w)) * 8) / 16
    maps[4] = 0.375                                                           # constant: every neighbour pair meets at a half integer
    maps[5, ::7, ::5] += 3.0                                                  # spikes: their parabolas own positions far beyond the window
    maps[6] = rng.standard_normal((h, w)).astype(np.float32)                  # noise
    maps[7, h // 2, w // 3] = np.nan                                           # non-finite samples: their lines go through the reference's
    maps[7, h // 4, w - 1] = -np.inf                                           # two loops literally (env::envelope_literal)
    maps[8, h // 3, w // 2] = np.inf
    defw = np.stack([rng.uniform(0.01, 0.02, n), rng.uniform(-0.02, 0.02, n), rng.uniform(0.01, 0.02, n), rng.uniform(-0.02, 0.02, n)], axis=1).astype(np.float32)
    defw[2] = (0.0625, 0.0, 0.125, -0.125)
    defw[3] = (0.03125, 0.0625, 0.5, 0.0)
    defw[4] = (0.0625, 0.0, 0.0625, 0.0625)
    anchors = np.stack([rng.integers(-3, 4, n), rng.integers(-2, 6, n)], axis=1).astype(np.int32)
    anchors[0] = (5, -5)
    anchors[9] = (-5, 5)
    L = oracle_lib.lib()
    ref = []
    for i in range(n):
        o, x, y = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
        L.orc_dt2d_f32(maps[i].reshape(-1), h, w, defw[i], int(anchors[i, 0]), int(anchors[i, 1]), 0, o.reshape(-1), x.reshape(-1), y.reshape(-1))
        ref.append((o, x, y))
    plan = Dt2dPlan(n, h, w, defw, anchors, 4)
    d_in = torch.from_numpy(maps).cuda()
    d_out = torch.empty_like(d_in)
    d_ix = torch.empty((n, h, w), dtype=torch.int16, device="cuda")
    d_iy = torch.empty_like(d_ix)
    for steps in (seg, 0, seg):                                               # segmented, one lane per line, segmented again (counters left clean)
        plan.set_segment(steps)
        d_out.zero_(); d_ix.zero_(); d_iy.zero_()
        plan.run(d_in.data_ptr(), d_out.data_ptr(), d_ix.data_ptr(), d_iy.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out, ix, iy = d_out.cpu().numpy(), d_ix.cpu().numpy().view(np.uint16), d_iy.cpu().numpy().view(np.uint16)
        for i in range(n):
            o, x, y = ref[i]
            assert np.array_equal(out[i], o, equal_nan=True), (i, steps)
            assert np.array_equal(ix[i].astype(np.int32), x) and np.array_equal(iy[i].astype(np.int32), y), (i, steps)
    assert plan.replayed() > 0
    plan.close()


@pytest.mark.parametrize("impl", [1, 3, 4])
def test_dt2d_non_finite_samples_equal_the_oracle(impl):
    """NaN and +-inf samples through the streaming kernels (eager and lagged-scan emission) and the windowed transform: their lines are
    redone with the reference's two loops, literally, so the result is whatever the reference's comparisons make of such a sample
    (the oracle is pinned to the compiled reference source on these inputs: tests/test_oracle_ref.py)."""
    import torch
    from partsbaseddetector_b200 import Dt2dPlan
    rng = np.random.default_rng(4711 + impl)
    n, h, w = 4, 45, 170
    maps = (rng.standard_normal((n, h, w)) * 0.3).astype(np.float32)
    for i in range(n):
        for k in range(5):
            maps[i, rng.integers(0, h), rng.integers(0, w)] = (np.nan, np.inf, -np.inf)[(i + k) % 3]
    maps[1, 0, 0] = np.nan; maps[2, h - 1, w - 1] = np.nan
    defw = np.stack([rng.uniform(0.01, 0.05, n), rng.uniform(-0.02, 0.02, n), rng.uniform(0.01, 0.05, n), rng.uniform(-0.02, 0.02, n)], axis=1).astype(np.float32)
    anchors = np.stack([rng.integers(-3, 4, n), rng.integers(-2, 6, n)], axis=1).astype(np.int32)
    plan = Dt2dPlan(n, h, w, defw, anchors, impl)
    d_in = torch.from_numpy(maps).cuda()
    d_out = torch.empty_like(d_in)
    d_ix = torch.empty((n, h, w), dtype=torch.int16, device="cuda")
    d_iy = torch.empty_like(d_ix)
    plan.run(d_in.data_ptr(), d_out.data_ptr(), d_ix.data_ptr(), d_iy.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out, ix, iy = d_out.cpu().numpy(), d_ix.cpu().numpy().view(np.uint16), d_iy.cpu().numpy().view(np.uint16)
    L = oracle_lib.lib()
    for i in range(n):
        o, x, y = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
        L.orc_dt2d_f32(maps[i].reshape(-1), h, w, defw[i], int(anchors[i, 0]), int(anchors[i, 1]), 0, o.reshape(-1), x.reshape(-1), y.reshape(-1))
        assert np.array_equal(out[i], o, equal_nan=True), i
        assert np.array_equal(ix[i].astype(np.int32), x) and np.array_equal(iy[i].astype(np.int32), y), i
    plan.close()


def test_dt2d_rejects_bad_arguments():
    with pytest.raises(PbdError):
        dt2d(np.zeros((4, 4), np.float32), [0.0, 0.0, 0.01, 0.0], [0, 0])       # a = -w0 must be < 0


# ------------------------------------------------------------------------------------------ DP + backtrack
@pytest.mark.parametrize("name,shape", [("Person_26parts", (120, 160)), ("Willowcoffee_5parts", (144, 200)),
                                        ("Face_frontal_sparse", (120, 160)), ("Person_8parts", (200, 260))])
@pytest.mark.parametrize("mode", [0, 1])
def test_dp_and_candidates_bitexact(name, shape, mode):
    fm = load_flat(name)
    img = synth_frame(21, *shape)
    d, O = detector(name), oracle(name)
    O.set_backptr_mode(mode)
    O.run(img, 1, 3)
    thr = lowered_threshold(O)
    O.set_thresh(thr)
    O.run(None, 4, 4)
    d.set_option("backptr", mode)
    d.set_option("thresh", thr)
    cands = d.detect(img)
    nl = O.nlevels()
    for l in range(nl):
        for c in range(len(fm.comps)):
            assert np.array_equal(d.rootv(0, l, c), O.rootv(l, c)), ("rootv", l, c)
            assert np.array_equal(d.rooti(0, l, c), O.rooti(l, c)), ("rooti", l, c)
    for l in (0, nl - 1):
        for c in range(len(fm.comps)):
            for p in range(1, len(fm.comps[c])):
                for pm in range(fm.nmix(c, fm.comps[c][p].parentid)):
                    g, o = d.backptr(0, l, c, p, pm), O.backptr(l, c, p, pm)
                    for a, b, what in zip(g, o, ("Ix", "Iy", "Ik")):
                        assert np.array_equal(a, b), (what, l, c, p, pm)
    oc = O.candidates()
    assert len(cands) == len(oc) > 0
    for g, o in zip(cands, oc):                                   # same deterministic order as the reference
        assert (g.frame, g.level, g.component()) == (0, o["level"], o["component"])
        assert np.array_equal(g.x, o["x"]) and np.array_equal(g.y, o["y"]) and np.array_equal(g.m, o["m"])
        assert np.array_equal(g.parts(), o["rects"])
        assert g.score() == o["score"] and np.all(g.confidence()[1:] == 0)


def test_committed_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "oracle_golden.npz"))
    d = detector("Person_26parts")
    d.set_option("thresh", float(g["p26_thresh"]))
    cands = d.detect(synth_frame(7, 120, 160))
    assert d.nscales() == int(g["p26_nlevels"])
    assert np.array_equal(d.pyramid_image(0, 2), g["p26_image2"])
    assert np.array_equal(d.features(0, 0), g["p26_feat0"])
    assert np.array_equal(d.response(0, 0, 17), g["p26_resp0_f17"])
    assert np.array_equal(d.rootv(0, 0), g["p26_rootv0"]) and np.array_equal(d.rooti(0, 0), g["p26_rooti0"])
    ix, iy, ik = d.backptr(0, 0, 0, 3, 2)
    assert np.array_equal(ix, g["p26_ix_p3m2"]) and np.array_equal(iy, g["p26_iy_p3m2"]) and np.array_equal(ik, g["p26_ik_p3m2"])
    got = np.array([[c.level] + list(c.x) + list(c.y) + list(c.m) for c in cands], np.int32)
    assert np.array_equal(got, g["p26_cand_xyms"])
    assert np.array_equal(np.array([c.score() for c in cands], np.float32), g["p26_cand_scores"])
    assert np.array_equal(np.array([c.parts() for c in cands], np.int32), g["p26_cand_rects"])


def test_cuda_path_equals_vectors_from_the_reference_sources():
    """tests/golden/ref_golden.npz holds outputs of the reference's OWN compiled sources (HOGFeatures.cpp, DistanceTransform.hpp,
    DynamicProgram.cpp, nms.cpp through oracle/_ref; generator: tests/golden/make_golden.py).  The CUDA path reproduces them bit for bit:
    HOG pyramid features and scales, 2-D distance transforms, DP root maps / back-pointers / candidates, root-map NMS."""
    g = np.load(os.path.join(GOLDEN, "ref_golden.npz"))
    fm = load_flat("Person_26parts")
    d = detector("Person_26parts")
    d.pyramid(synth_frame(11, 120, 160))
    assert d.nscales() == int(g["hog_nlevels"])
    for l in range(d.nscales()):
        assert np.array_equal(d.features(0, l), g["hog_feat%d" % l]) and d.level_info(l)["scale"] == g["hog_scales"][l]
    out, ix, iy = dt2d(g["dt_in"], g["dt_defw"], g["dt_anchor"], 0)
    assert np.array_equal(out, g["dt_out"]) and np.array_equal(ix, g["dt_ix"]) and np.array_equal(iy, g["dt_iy"])
    ohow = [tuple(int(v) for v in r) for r in g["dp_ohow"]]
    thr = float(g["dp_thresh"])
    d.set_levels(1, ohow, g["dp_scales"])
    for l, shp in enumerate(ohow):
        for f in range(fm.nfilters()):
            d.set_response(0, l, f, (np.random.default_rng(1000 * l + f).standard_normal(shp) * 0.3).astype(np.float32))
    d.set_option("thresh", thr)
    d.min()
    for l in range(2):
        assert np.array_equal(d.rootv(0, l), g["dp_rootv%d" % l]) and np.array_equal(d.rooti(0, l), g["dp_rooti%d" % l])
    for (p, pm) in ((1, 0), (3, 2), (14, 4), (25, 1)):
        assert np.array_equal(np.stack(d.backptr(0, 0, 0, p, pm)), g["dp_bp_p%d_m%d" % (p, pm)])
    cands = sorted(d.argmin(), key=lambda c: (float(c.score()), np.ascontiguousarray(c.parts(), np.int32).tobytes()))
    assert np.array_equal(np.stack([c.parts() for c in cands]), g["dp_cand_rects"])
    assert np.array_equal(np.array([c.score() for c in cands], np.float32), g["dp_cand_scores"])
    d.set_option("root_nms", 2)
    kept = [c for c in d.argmin() if c.level == 0]
    mask = (g["dp_rootv0"] > np.float32(thr)) & (g["dp_rootnms2_level0"] > 0)
    assert len(kept) == int(mask.sum()) > 0 and all(mask[c.y[0], c.x[0]] for c in kept)
    d.set_option("root_nms", 0)


def test_cuda_stages_feed_the_reference_dynamic_program():
    """Mixed pipeline, as a host that swaps single stages would run it: image pyramid + HOG + part responses on the GPU (the IFeatures /
    IConvolutionEngine stages), then the REFERENCE'S OWN DynamicProgram<float>::min / argmin (src/DynamicProgram.cpp compiled unmodified
    in oracle/_ref) on those responses.  Its root maps, back-pointers and candidates equal the all-CUDA path's bit for bit."""
    import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref/libpbd_ref.so not built")
    fm = load_flat("Person_26parts")
    d = detector("Person_26parts")
    img = synth_frame(123, 120, 160)
    d.set_option("thresh", 1e9)
    d.detect(img)
    nl = d.nscales()
    rv = np.sort(np.concatenate([d.rootv(0, l).ravel() for l in range(nl)]))
    thr = float(0.5 * (float(rv[-41]) + float(rv[-40])))
    d.set_option("thresh", thr)
    cands = d.detect(img)
    R = ref_lib.RefDP(fm, 32)
    R.set_levels([(d.level_info(l)["oh"], d.level_info(l)["ow"]) for l in range(nl)], np.array(d.scales(), np.float32))
    for l in range(nl):
        for f in range(fm.nfilters()):
            R.set_response(l, f, d.response(0, l, f))
    assert R.run(thr) == len(cands) == 40
    for l in range(nl):
        v, i = R.root(l)
        assert np.array_equal(v.astype(np.float32), d.rootv(0, l)) and np.array_equal(i, d.rooti(0, l))
    for (p, pm) in ((2, 3), (13, 0), (25, 4)):
        assert all(np.array_equal(a, b) for a, b in zip(R.backptr(0, 0, p, pm), d.backptr(0, 0, 0, p, pm)))
    key = lambda rects, score: (float(score), np.ascontiguousarray(rects, np.int32).tobytes())
    assert sorted(key(r, c[0]) for _, r, c in R.candidates()) == sorted(key(c.parts(), c.score()) for c in cands)


def test_fast_mode_integer_outputs_and_score_tolerance():
    # fused multiply-add responses: scores within 1e-4 relative (north_star), integer outputs expected identical
    name = "Person_26parts"
    img = synth_frame(33, 240, 320)
    d, O = detector(name), oracle(name)
    O.run(img, 1, 3)
    thr = lowered_threshold(O, 200)
    O.set_thresh(thr + 1e-4)          # keep clear of the threshold so the candidate SET cannot flip on a 1e-7 score change
    O.run(None, 4, 4)
    d.set_option("exact", 0)
    d.set_option("thresh", thr + 1e-4)
    cands = d.detect(img)
    oc = O.candidates()
    gk = {(c.level, int(c.y[0]), int(c.x[0])): c for c in cands}
    ok = {(c["level"], int(c["y"][0]), int(c["x"][0])): c for c in oc}
    common = set(gk) & set(ok)
    assert len(common) >= 0.98 * len(ok)
    same = 0
    for k in common:
        g, o = gk[k], ok[k]
        assert abs(g.score() - o["score"]) <= 1e-4 * abs(o["score"])
        same += int(np.array_equal(g.x, o["x"]) and np.array_equal(g.y, o["y"]) and np.array_equal(g.m, o["m"]))
    assert same >= 0.98 * len(common)


# ------------------------------------------------------------------------------------------ batches, full size, properties
def test_batch_equals_single_frames_vga():
    name = "Person_26parts"
    frames = synth_frames(3, 480, 640, start=100)
    d = detector(name)
    O = oracle(name)
    O.run(frames[1], 1, 3)
    thr = lowered_threshold(O, 80)
    d.set_option("thresh", thr)
    batch = d.detect(frames)
    rv_batch = [d.rootv(f, 0).copy() for f in range(3)]
    singles = []
    for f in range(3):
        c = d.detect(frames[f])
        assert np.array_equal(d.rootv(0, 0), rv_batch[f])
        singles += [(f, k.level, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score())) for k in c]
    assert [(k.frame, k.level, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score())) for k in batch] == singles
    # frame 1 of the batch against the oracle at full VGA size (14 levels)
    O.set_thresh(thr)
    O.run(None, 4, 4)
    oc = O.candidates()
    b1 = [k for k in batch if k.frame == 1]
    assert len(b1) == len(oc) > 0
    for g, o in zip(b1, oc):
        assert g.level == o["level"] and np.array_equal(g.x, o["x"]) and np.array_equal(g.y, o["y"]) and np.array_equal(g.m, o["m"])
        assert g.score() == o["score"] and np.array_equal(g.parts(), o["rects"])


@pytest.mark.parametrize("nframes", [8, 11, 17])
def test_dp_frame_groups_on_concurrent_streams_do_not_change_results(nframes):
    # the DP stage runs the batch as dp_streams groups of frames on concurrent streams (engine.cu run_dp_min): every output of
    # every frame must equal the single-stream result bit for bit, and frame 0 / the last frame must equal the oracle
    name = "Person_26parts"
    frames = synth_frames(nframes, 120, 168, start=300)
    d, O = detector(name), oracle(name)
    O.run(frames[nframes - 1], 1, 3)
    thr = lowered_threshold(O, 40)
    d.set_option("thresh", thr)
    ref = None
    for ns in (1, 2, 3, 4):
        d.set_option("dp_streams", ns)
        assert d.get_option("dp_streams") == ns
        cands = d.detect(frames)
        got = {"cands": [(k.frame, k.level, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score())) for k in cands],
               "rootv": [d.rootv(f, l).copy() for f in range(nframes) for l in range(d.nscales())],
               "rooti": [d.rooti(f, l).copy() for f in range(nframes) for l in range(d.nscales())],
               "bp": [np.stack(d.backptr(f, 0, 0, p, m)) for f in (0, nframes // 2, nframes - 1) for p, m in ((3, 2), (25, 0), (12, 4))]}
        if ref is None:
            ref = got
            continue
        assert got["cands"] == ref["cands"] and len(got["cands"]) > 0
        for key in ("rootv", "rooti", "bp"):
            assert all(np.array_equal(a, b) for a, b in zip(got[key], ref[key])), (ns, key)
    for l in range(O.nlevels()):
        assert np.array_equal(d.rootv(nframes - 1, l), O.rootv(l))
    d.set_option("dp_streams", 2)
    with pytest.raises(PbdError):
        d.set_option("dp_streams", 0)


@pytest.mark.parametrize("name,shape,keep", [("Person_26parts", (240, 320), 400), ("Face_99filters", (160, 200), 120), ("Person_8parts", (203, 177), 250)])
def test_device_sort_and_nms_equal_host_sort_and_nms(name, shape, keep):
    # option nms_overlap >= 0: Candidate::sort + Candidate::nonMaximaSuppression run on the device per frame and only the survivors
    # are downloaded; the host versions (pbd_candidates_sort / pbd_candidates_nms, parity-tested against the oracle's restatement in
    # test_host_logic.py) applied to the raw candidates of each frame must give the same list
    from partsbaseddetector_b200 import Candidate
    frames = synth_frames(5, shape[0], shape[1], start=900)
    frames[3] = 0                                           # a flat frame: every score ties (canonical order decides)
    d, O = detector(name), oracle(name)
    if name == "Face_99filters":
        d.set_option("max_levels", 1)
        O.set_max_levels(1)
    O.run(frames[0], 1, 3)
    thr = lowered_threshold(O, keep)
    d.set_option("thresh", thr)
    d.set_option("nms_overlap", -1)
    raw = [list(d.detect(frames[f])) for f in range(5)]
    assert sum(len(r) for r in raw) > 50
    sig = lambda k: (k.level, k.component_, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score()), k.parts().tobytes())
    for ov in (0.0, 0.15, 0.6):
        d.set_option("nms_overlap", ov)
        got = list(d.detect(frames))
        assert d.get_option("nms_overlap") == pytest.approx(ov)
        for f in range(5):
            want = Candidate.nonMaximaSuppression(frames[f], Candidate.sort(list(raw[f])), ov)
            mine = [k for k in got if k.frame == f]
            assert [sig(k) for k in mine] == [sig(k) for k in want], (ov, f, len(mine), len(want))
        assert [k.frame for k in got] == sorted(k.frame for k in got)
        # the streaming API goes through the same path
        assert [sig(k) for k in d.collect_ticket(d.submit(frames))] == [sig(k) for k in got]
    d.set_option("nms_overlap", -1)
    assert len(d.detect(frames)) == sum(len(r) for r in raw)


def test_determinism_and_stage_api_equivalence():
    d = detector("Person_26parts")
    frames = synth_frames(2, 240, 320, start=7)
    d.set_option("thresh", -1.2)
    a = d.detect(frames)
    b = d.detect(frames)
    assert [(k.frame, k.level, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score())) for k in a] == \
           [(k.frame, k.level, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score())) for k in b]
    d.pyramid(frames)
    d.pdf()
    d.min()
    c = d.argmin()
    assert [(k.frame, k.level, tuple(k.x), tuple(k.y), tuple(k.m)) for k in a] == [(k.frame, k.level, tuple(k.x), tuple(k.y), tuple(k.m)) for k in c]


def test_dt_kernel_variants_give_identical_results():
    """option dt_variant: 0 = every break point in double (default), 1 = certified fp32 break points, 2 = lagged-scan
    emission, 3 = windowed certified transform: root maps, back-pointers and candidates are the same bits (and equal the oracle's,
    tested elsewhere for the default)."""
    img = np.stack([synth_frame(900 + i, 150, 210) for i in range(3)])
    img[2, 40:90, 30:120] = 17                                   # a flat patch: exact ties between break points
    outs = []
    for variant in (0, 1, 2, 3):
        d = detector("Person_26parts")
        d.set_option("dt_variant", variant)
        d.set_option("thresh", -1.25)
        c = d.detect(img)
        maps = [d.rootv(f, l) for f in range(3) for l in range(d.nscales())] + [np.stack(d.backptr(2, 0, 0, p, 1)) for p in (1, 7, 25)]
        outs.append((maps, [(a.frame, a.level, a.x.tolist(), a.y.tolist(), a.m.tolist(), float(a.score())) for a in c]))
        d.set_option("dt_variant", 3)                            # the helper's detectors are shared between tests (3 = the default)
    assert len(outs[0][1]) > 20
    for other in outs[1:]:
        assert other[1] == outs[0][1]
        assert all(np.array_equal(a, b) for a, b in zip(other[0], outs[0][0]))


@pytest.mark.parametrize("name", ["Person_26parts", "Person_8parts", "Face_99filters", "Willowcoffee_5parts", "Face_frontal_sparse"])
@pytest.mark.parametrize("backptr", [0, 1])
def test_windowed_transform_equals_the_stack_algorithm(name, backptr):
    """dt_variant 3 (dt_window.cu): maps whose anchors fit the window go through the certified position-parallel kernels, refused
    lines through the replay kernel, the other maps (large anchors: the face models, Person_8parts) keep dt_pass -- a mixed wave.
    Smooth frames, a noise frame (many refusals), a constant frame (every break point a tie) and a frame with flat patches; root
    maps, arg-max maps of every part and candidates must be the bits of variant 0; and they equal the oracle on the first frame."""
    rng = np.random.default_rng(77)
    frames = np.stack([synth_frame(500 + i, 140, 200) for i in range(6)])
    frames[1] = rng.integers(0, 256, frames[1].shape, dtype=np.uint8)
    frames[2] = 93
    frames[3, 30:100, 20:150] = 200
    fm = load_flat(name)
    res = {}
    # (variant, dt_segment): the stack kernel, the windowed walk with one lane per line, and the SEGMENTED walk (lines of the 33 x 48
    # cell level cut into 3 / 2 segments of 22 / 38 positions, as the detector does by itself for launches that cannot fill the GPU)
    for variant, seg in ((0, 0), (3, 0), (3, 32), (3, 48), (3, -1)):
        d = detector(name)
        d.set_option("dt_variant", variant); d.set_option("backptr", backptr); d.set_option("dt_segment", seg)
        if variant == 3:
            d.get_option("dt_replayed_lines")                        # reset the counter
        c = d.detect(frames)
        maps = []
        for f in range(6):
            for l in range(d.nscales()):
                for comp in range(len(fm.comps)):
                    maps.append(d.rootv(f, l, comp)); maps.append(d.rooti(f, l, comp))
        for f in (0, 1, 2, 3):
            for p in range(1, len(fm.comps[0])):
                npm = len(fm.comps[0][fm.comps[0][p].parentid].filterid)
                maps.append(np.stack(d.backptr(f, 0, 0, p, npm - 1)))
        res[(variant, seg)] = (maps, [(a.frame, a.level, a.x.tolist(), a.y.tolist(), a.m.tolist(), float(a.score())) for a in c])
        if (variant, seg) == (3, 0):
            replayed = d.get_option("dt_replayed_lines")
        d.set_option("dt_variant", 3); d.set_option("backptr", 0); d.set_option("dt_segment", -1)
    for key in ((3, 0), (3, 32), (3, 48), (3, -1)):
        assert len(res[(0, 0)][0]) == len(res[key][0])
        for i, (a, b) in enumerate(zip(res[(0, 0)][0], res[key][0])):
            assert np.array_equal(a, b), (name, key, i)
        assert res[(0, 0)][1] == res[key][1], key
    if name in ("Person_8parts", "Face_99filters", "Face_frontal_sparse"):
        assert replayed > 0                                          # anchors beyond the window: those maps' lines all take the replay path


def test_segmented_walk_single_vga_frame_equals_the_oracle():
    """A single VGA frame: the DP launches cannot fill the GPU, so dt_variant 3 cuts the lines into segments by itself (dt_segment -1).
    Root maps, every arg-max map of the first and a middle level and the candidates must equal the CPU oracle; the forced segment
    lengths give the same bits."""
    name = "Person_26parts"
    fm = load_flat(name)
    img = synth_frame(4242)
    O = oracle(name)
    O.run(img, 1, 3)
    thr = lowered_threshold(O)
    O.set_thresh(thr)
    O.run(None, 4, 4)
    oc = O.candidates()
    assert len(oc) > 0
    for seg in (-1, 32, 64, 0):
        d = detector(name)
        d.set_option("thresh", thr); d.set_option("dt_segment", seg)
        cands = d.detect(img)
        for l in range(O.nlevels()):
            assert np.array_equal(d.rootv(0, l, 0), O.rootv(l, 0)), (seg, l)
            assert np.array_equal(d.rooti(0, l, 0), O.rooti(l, 0)), (seg, l)
        for l in (0, 5):
            for p in range(1, len(fm.comps[0])):
                for pm in range(fm.nmix(0, fm.comps[0][p].parentid)):
                    for a, b, what in zip(d.backptr(0, l, 0, p, pm), O.backptr(l, 0, p, pm), ("Ix", "Iy", "Ik")):
                        assert np.array_equal(a, b), (seg, what, l, p, pm)
        assert len(cands) == len(oc)
        for g, o in zip(cands, oc):
            assert g.level == o["level"] and np.array_equal(g.x, o["x"]) and np.array_equal(g.y, o["y"]) and np.array_equal(g.m, o["m"])
            assert g.score() == o["score"]
        d.set_option("dt_segment", -1)


def test_cuda_graph_replay_equals_eager_launches():
    """option graph=1: the second enqueue with the same frames pointer / geometry / options captures the launch sequence (frame groups
    of the DP on forked streams included), later ones replay it; results equal the eager path, a changed option re-captures."""
    import torch
    frames = np.stack([synth_frame(300 + i, 120, 160) for i in range(8)])
    dev = torch.from_numpy(frames).cuda()
    d = PartsBasedDetector()
    d.distributeModel(Model.load_bin(golden_model_path("Person_26parts")))
    d.set_option("thresh", -1.3)
    ref = d.detect(frames)
    assert len(ref) > 20
    d.set_option("graph", 1)
    for it in range(4):                                   # eager (warm-up), capture + launch, replay, replay
        d.enqueue_device(dev.data_ptr(), 8, 120, 160, 3)
        got = d.collect()
        assert len(got) == len(ref), it
        for a, b in zip(got, ref):
            assert a.frame == b.frame and a.level == b.level and np.array_equal(a.x, b.x) and np.array_equal(a.y, b.y) and np.array_equal(a.m, b.m)
            assert a.score() == b.score() and np.array_equal(a.parts(), b.parts())
    n_before = d.launch_count()
    d.enqueue_device(dev.data_ptr(), 8, 120, 160, 3)
    assert d.launch_count() > n_before                    # replayed kernels are counted
    d.set_option("thresh", -1.2)                          # an option the captured sequence depends on: eager again, then a new graph
    ref2 = d.detect(frames)
    for it in range(3):
        d.enqueue_device(dev.data_ptr(), 8, 120, 160, 3)
        got = d.collect()
        assert len(got) == len(ref2) and all(a.score() == b.score() and np.array_equal(a.parts(), b.parts()) for a, b in zip(got, ref2))
    other = torch.from_numpy(frames[::-1].copy()).cuda()  # another frames pointer: must not replay the old graph
    d.enqueue_device(other.data_ptr(), 8, 120, 160, 3)
    got = d.collect()
    assert [c.frame for c in got] != [] and sorted(float(c.score()) for c in got) == sorted(float(c.score()) for c in ref2)
    d.close()


@pytest.mark.parametrize("sz", [1, 3])
def test_root_map_nms_option_equals_oracle(sz):
    """option root_nms = sz: only the strict local maxima of every root map (nonMaximaSuppression of reference src/nms.cpp:84-129, window
    sz) above the threshold are backtracked; equals the oracle's restatement of that function applied to the oracle's root maps."""
    d, O = detector("Person_26parts"), oracle("Person_26parts")
    img = synth_frame(77, 200, 280)
    O.run(img, 1, 3)
    thr = lowered_threshold(O, 400)
    O.set_thresh(thr)
    O.run(None, 4, 4)
    L = oracle_lib.lib()
    keep = {}
    for l in range(O.nlevels()):
        rv = O.rootv(l)
        k = np.empty(rv.shape, np.uint8)
        L.orc_rootmap_nms(np.ascontiguousarray(rv).reshape(-1), rv.shape[0], rv.shape[1], sz, None, k.reshape(-1))
        keep[l] = k
    want = [o for o in O.candidates() if keep[o["level"]][o["y"][0], o["x"][0]]]
    d.set_option("thresh", thr)
    d.set_option("root_nms", sz)
    got = d.detect(img)
    assert 0 < len(want) < len(O.candidates()) and len(got) == len(want)
    for a, b in zip(got, want):
        assert a.level == b["level"] and np.array_equal(a.x, b["x"]) and np.array_equal(a.y, b["y"]) and np.array_equal(a.m, b["m"])
        assert a.score() == b["score"] and np.array_equal(a.parts(), b["rects"])
    d.set_option("root_nms", 0)
    assert len(d.detect(img)) == len(O.candidates())


def test_max_levels_and_candidate_sort():
    d, O = detector("Person_26parts"), oracle("Person_26parts")
    img = synth_frame(9, 240, 320)
    d.set_option("max_levels", 4)
    O.set_max_levels(4)
    O.run(img, 1, 3)
    thr = lowered_threshold(O, 30)
    O.set_thresh(thr)
    O.run(None, 4, 4)
    d.set_option("thresh", thr)
    cands = d.detect(img)
    assert d.nscales() == 4 and len(cands) == len(O.candidates())
    from partsbaseddetector_b200 import Candidate
    cands = Candidate.sort(list(cands))
    s = [float(c.score()) for c in cands]
    assert s == sorted(s, reverse=True)


def test_errors_and_state_machine():
    d = detector("Person_26parts")
    with pytest.raises(PbdError):
        d.detect(np.zeros((8, 8, 3), np.uint8))                 # smaller than one pyramid level
    with pytest.raises(PbdError):
        d.detect(np.zeros((64, 64, 3), np.float32))             # unsupported depth
    d.pyramid(synth_frame(1, 96, 128))
    with pytest.raises(PbdError):
        d.min()                                                 # pdf has not run
    d.set_option("max_candidates", 4)
    d.set_option("thresh", -100.0)
    with pytest.raises(PbdError):
        d.detect(synth_frame(1, 96, 128))                       # candidate buffer overflow is reported, not truncated
    d.set_option("max_candidates", 65536)
    with pytest.raises(PbdError):
        d.set_option("nonsense", 1)


def test_cpp_adapter_demo_matches_python(tmp_path):
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "demo")
    libdir = os.path.join(ROOT, "partsbaseddetector_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "demo.cpp"),
                           "-L" + libdir, "-lpbd_b200", "-Wl,-rpath," + libdir, "-o", exe])
    img = synth_frame(3, 120, 160)
    ppm = tmp_path / "f.ppm"
    ppm.write_bytes(b"P6\n160 120\n255\n" + np.ascontiguousarray(img[:, :, ::-1]).tobytes())
    d = detector("Person_26parts")
    d.set_option("thresh", -1.2)
    cands = d.detect(img)
    r = subprocess.run([exe, os.path.join(GOLDEN, "Person_26parts.pbdm"), str(ppm), "-1.2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "Number of candidates: %d" % len(cands) in r.stdout
    best = max(cands, key=lambda c: float(c.score()))
    assert ("best: score %.6f level %d" % (float(best.score()), best.level)) in r.stdout


# ------------------------------------------------------------------------------------------ BASELINE.json full sizes
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_stage_plugins_match_oracle(tmp_path, precision):
    """CudaHOGFeatures : IFeatures and CudaConvolutionEngine : IConvolutionEngine (include/pbd_b200_plugins.hpp) driven through the
    reference's signatures by examples/plugin_demo.cpp: pyramid() feature Mats and pdf() response Mats of every level equal the
    oracle's bit for bit (T = double: the fp32 results widened, as documented)."""
    import subprocess
    from conftest import ROOT
    libdir = os.path.join(ROOT, "partsbaseddetector_b200")
    exe = str(tmp_path / "plugin_demo")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "ref_shim"),
                           os.path.join(ROOT, "examples", "plugin_demo.cpp"), "-L" + libdir, "-lpbd_b200", "-Wl,-rpath," + libdir, "-o", exe])
    img = synth_frame(61, 150, 200)
    (tmp_path / "f.raw").write_bytes(img.tobytes())
    out = tmp_path / "o.bin"
    r = subprocess.run([exe, golden_model_path("Person_26parts"), str(tmp_path / "f.raw"), "150", "200", "3", str(out)] + (["f64"] if precision == "f64" else []),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    O = oracle("Person_26parts")
    O.run(img, 1, 2)
    buf = out.read_bytes()
    n, nf, es = np.frombuffer(buf, np.int32, 3, 0)
    dt = np.float64 if es == 8 else np.float32
    assert n == O.nlevels() and nf == 138 and es == (8 if precision == "f64" else 4)
    off = 12
    for l in range(n):
        oh, owf = np.frombuffer(buf, np.int32, 2, off)
        scale = np.frombuffer(buf, np.float32, 1, off + 8)[0]
        off += 12
        li = O.level_info(l)
        assert oh == li["oh"] and owf == li["ow"] * 32 and scale == li["scale"]
        feat = np.frombuffer(buf, dt, oh * owf, off).reshape(oh, owf // 32, 32)
        off += oh * owf * es
        assert np.array_equal(feat, O.features(l).astype(dt)), l
        for f in range(nf):
            resp = np.frombuffer(buf, dt, oh * owf // 32, off).reshape(oh, owf // 32)
            off += oh * (owf // 32) * es
            assert np.array_equal(resp, O.response(l, f).astype(dt)), (l, f)
    assert off == len(buf)


def test_config4_1080p_ten_levels_bitexact():
    """config 4: person model, 1920x1080, first 10 pyramid levels (340 400 cells)."""
    name = "Person_26parts"
    img = synth_frame(77, 1080, 1920)
    d, O = detector(name), oracle(name)
    d.set_option("max_levels", 10)
    O.set_max_levels(10)
    O.run(img, 1, 3)
    thr = lowered_threshold(O, 120)
    O.set_thresh(thr)
    O.run(None, 4, 4)
    d.set_option("thresh", thr)
    cands = d.detect(img)
    assert d.nscales() == 10 and d.level_info(0)["ow"] == 478 and d.level_info(0)["oh"] == 268
    for l in (0, 4, 9):
        assert np.array_equal(d.rootv(0, l), O.rootv(l)) and np.array_equal(d.rooti(0, l), O.rooti(l))
    g, o = d.backptr(0, 0, 0, 13, 1), O.backptr(0, 0, 13, 1)
    assert all(np.array_equal(a, b) for a, b in zip(g, o))
    oc = O.candidates()
    assert len(cands) == len(oc) > 0
    for a, b in zip(cands, oc):
        assert a.level == b["level"] and np.array_equal(a.x, b["x"]) and np.array_equal(a.y, b["y"]) and np.array_equal(a.m, b["m"])
        assert a.score() == b["score"] and np.array_equal(a.parts(), b["rects"])


def test_config5_dt_4096_bitexact_and_properties():
    """config 5 (one of the 156 maps): 4096x4096 score map through the standalone DT, against the oracle, plus
    size-independent properties: out >= in + penalty at the anchor, and the exact back-pointers attain the output."""
    h = w = 4096
    m = synth_score_map(0, h, w)
    defw = np.array([[0.013, -0.004, 0.017, 0.011]], np.float32)
    anchor = np.array([[2, -1]], np.int32)
    out, ix, iy = dt2d(m, defw, anchor, 1)
    o, x, y = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
    oracle_lib.lib().orc_dt2d_f32(m.reshape(-1), h, w, defw[0], 2, -1, 1, o.reshape(-1), x.reshape(-1), y.reshape(-1))
    assert np.array_equal(out[0], o) and np.array_equal(ix[0], x) and np.array_equal(iy[0], y)
    a0, b0, a1, b1 = (-float(v) for v in defw[0])
    yy, xx = np.mgrid[0:h:97, 0:w:89]                       # sampled cells
    dx, dy = (xx + 2) - ix[0][yy, xx], (yy - 1) - iy[0][yy, xx]
    att = m[iy[0][yy, xx], ix[0][yy, xx]] + a0 * dx * dx + b0 * dx + a1 * dy * dy + b1 * dy
    assert np.allclose(att, out[0][yy, xx], rtol=1e-5, atol=1e-5)
    inside = (xx + 2 < w) & (yy - 1 >= 0)
    assert np.all(out[0][yy, xx][inside] >= m[np.clip(yy - 1, 0, h - 1), np.clip(xx + 2, 0, w - 1)][inside] - 1e-6)


def test_config1_face_qvga_single_scale():
    """BASELINE.json config 1: face model (Face_99filters stands in for the missing Face_68parts named by
    conf/config_face.by_parts:31), one 320x240 frame, level 0 only: 13 components, 39/68 single-mixture parts, 99 shared filters."""
    name = "Face_99filters"
    fm = load_flat(name)
    img = synth_frame(31, 240, 320)
    d, O = detector(name), oracle(name)
    d.set_option("max_levels", 1)
    O.set_max_levels(1)
    O.run(img, 1, 3)
    thr = lowered_threshold(O, 40)
    O.set_thresh(thr)
    O.run(None, 4, 4)
    d.set_option("thresh", thr)
    cands = d.detect(img)
    assert d.nscales() == 1 and (d.level_info(0)["oh"], d.level_info(0)["ow"]) == (58, 78)
    for c in range(len(fm.comps)):
        assert np.array_equal(d.rootv(0, 0, c), O.rootv(0, c)) and np.array_equal(d.rooti(0, 0, c), O.rooti(0, c)), c
    for (c, p) in ((0, 1), (3, 20), (12, len(fm.comps[12]) - 1)):
        g, o = d.backptr(0, 0, c, p, 0), O.backptr(0, c, p, 0)
        assert all(np.array_equal(a, b) for a, b in zip(g, o)), (c, p)
    oc = O.candidates()
    assert len(cands) == len(oc) > 0
    for a, b in zip(cands, oc):
        assert (a.level, a.component()) == (b["level"], b["component"])
        assert np.array_equal(a.x, b["x"]) and np.array_equal(a.y, b["y"]) and np.array_equal(a.m, b["m"])
        assert a.score() == b["score"] and np.array_equal(a.parts(), b["rects"])


def test_large_batch_indexing():
    """70 frames (7 distinct, cycled) in one batch: every frame's root scores and candidates equal the single-frame run
    (exercises the frame strides of every buffer beyond the small batches used elsewhere)."""
    d = detector("Person_26parts")
    base = synth_frames(7, 144, 192, start=300)
    batch = np.ascontiguousarray(np.stack([base[i % 7] for i in range(70)]))
    d.set_option("thresh", 1e9)
    d.detect(base[0])
    rv = np.concatenate([d.rootv(0, l).ravel() for l in range(d.nscales())])
    d.set_option("thresh", float(np.sort(rv)[-30]))            # ~30 candidates per frame
    singles = []
    for i in range(7):
        c = d.detect(base[i])
        singles.append(([(k.level, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score())) for k in c], d.rootv(0, 0).copy(), d.rootv(0, d.nscales() - 1).copy()))
    allc = d.detect(batch)
    per_frame = {}
    for k in allc:
        per_frame.setdefault(k.frame, []).append((k.level, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score())))
    assert sum(len(s[0]) for s in singles) > 0
    for f in range(70):
        assert per_frame.get(f, []) == singles[f % 7][0], f
    for f in (0, 33, 69):
        assert np.array_equal(d.rootv(f, 0), singles[f % 7][1]) and np.array_equal(d.rootv(f, d.nscales() - 1), singles[f % 7][2])


def test_pipelined_submit_collect_equals_detect():
    """pbd_submit_batch_u8 / pbd_collect_ticket with two batches in flight give the same candidates as detect()."""
    d = detector("Person_26parts")
    batches = [np.ascontiguousarray(synth_frames(9, 144, 192, start=400 + 9 * i)) for i in range(5)]
    d.set_option("thresh", 1e9)
    d.detect(batches[0])
    rv = np.concatenate([d.rootv(0, l).ravel() for l in range(d.nscales())])
    d.set_option("thresh", float(np.sort(rv)[-25]))
    key = lambda c: [(k.frame, k.level, tuple(k.x), tuple(k.y), tuple(k.m), float(k.score())) for k in c]
    ref = [key(d.detect(b)) for b in batches]
    got = [key(c) for c in d.detect_stream(batches)]
    assert got == ref and sum(map(len, ref)) > 0
    t0 = d.submit(batches[0])
    t1 = d.submit(batches[1])
    with pytest.raises(PbdError):
        d.submit(batches[2])                      # only two batches may be in flight
    assert key(d.collect_ticket(t1)) == ref[1] and key(d.collect_ticket(t0)) == ref[0]
    with pytest.raises(PbdError):
        d.collect_ticket(t0)
    assert key(d.detect(batches[3])) == ref[3]     # the synchronous API still works afterwards


@pytest.mark.parametrize("kind", ["zeros", "saturated", "ramp", "checker", "noise"])
def test_degenerate_frames_end_to_end(kind):
    """Flat / saturated / periodic frames give constant or exactly repeating score maps: every distance transform is full of
    exact ties, the hardest case for reproducing the reference's break-point rounding.  Everything is compared bit for bit."""
    h, w = 132, 180
    yy, xx = np.mgrid[0:h, 0:w]
    img = {"zeros": np.zeros((h, w, 3), np.uint8),
           "saturated": np.full((h, w, 3), 255, np.uint8),
           "ramp": np.stack([(xx * 255 // (w - 1)).astype(np.uint8)] * 3, axis=-1),
           "checker": np.stack([(((xx // 8 + yy // 8) % 2) * 200).astype(np.uint8)] * 3, axis=-1),
           "noise": np.random.default_rng(9).integers(0, 256, (h, w, 3)).astype(np.uint8)}[kind]
    img = np.ascontiguousarray(img)
    name = "Person_26parts"
    d, O = detector(name), oracle(name)
    O.run(img, 1, 3)
    rv = np.concatenate([O.rootv(l).ravel() for l in range(O.nlevels())])
    thr = float(np.sort(np.unique(rv))[-min(5, np.unique(rv).size)]) - 1e-6     # a handful of distinct top scores (ties included)
    O.set_thresh(thr)
    O.run(None, 4, 4)
    d.set_option("thresh", thr)
    d.set_option("max_candidates", 1 << 20)
    cands = d.detect(img)
    for l in range(O.nlevels()):
        assert np.array_equal(d.rootv(0, l), O.rootv(l)) and np.array_equal(d.rooti(0, l), O.rooti(l)), l
    for p, pm in ((1, 0), (13, 2), (25, 4)):
        g, o = d.backptr(0, 0, 0, p, pm), O.backptr(0, 0, p, pm)
        assert all(np.array_equal(a, b) for a, b in zip(g, o)), (p, pm)
    oc = O.candidates()
    assert len(cands) == len(oc) > 0
    for a, b in zip(cands, oc):
        assert a.level == b["level"] and np.array_equal(a.x, b["x"]) and np.array_equal(a.y, b["y"]) and np.array_equal(a.m, b["m"])
        assert a.score() == b["score"]
    d.set_option("max_candidates", 65536)


def test_environment_defaults(monkeypatch):
    monkeypatch.setenv("PBD_EXACT", "0")
    monkeypatch.setenv("PBD_BACKPTR", "exact")
    monkeypatch.setenv("PBD_MAX_LEVELS", "3")
    d = PartsBasedDetector(device=0)
    d.distributeModel(Model.load_bin(os.path.join(GOLDEN, "Willowcoffee_5parts.pbdm")))
    assert (d.get_option("exact"), d.get_option("backptr"), d.get_option("max_levels")) == (0.0, 1.0, 3.0)
    d.pyramid(synth_frame(1, 200, 260))
    assert d.nscales() == 3
    d.close()
