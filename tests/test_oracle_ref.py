"""Pins the restated oracle (oracle/pbd_oracle.cpp) BIT FOR BIT to the reference's own sources, compiled unmodified from
/root/reference against a minimal cv:: stand-in (oracle/_ref/libpbd_ref.so, recipe: oracle/Makefile target `ref`):

  DistanceTransform<T>::compute   include/DistanceTransform.hpp:203-245 (computeRow :152-182)
  Math::reduceMax / reducePickIndex   include/Math.hpp:109-185
  HOGFeatures<T>::pyramid / features  src/HOGFeatures.cpp:95-151, 168-341
  DynamicProgram<T>::min / argmin     src/DynamicProgram.cpp:67-255 (+ Parts.hpp indirection)
  Candidate::sort / nonMaximaSuppression  include/Candidate.hpp:97-99, 277-304

What the stand-in supplies is data movement only (Mat headers, transposes, elementwise + / += / > in the Mat's depth, saturate_cast);
cv::resize / cv::pyrDown come from the oracle's restatements, which test_oracle_pins.py pins bit-exact to cv2.  Not covered: the
part-filter responses (the reference's convolution is a 4 kLoC copy of OpenCV's FilterEngine that needs OpenCV's internal headers)."""
import numpy as np
import pytest

import oracle_lib
import ref_lib
from conftest import load_flat
from partsbaseddetector_b200.synth import synth_frame, synth_score_map

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libpbd_ref.so not built (needs /root/reference)")


@pytest.mark.parametrize("h,w", [(1, 1), (1, 9), (7, 1), (13, 17), (58, 78), (118, 158), (200, 331)])
def test_distance_transform_equals_reference_source(h, w):
    rng = np.random.default_rng(h * 131 + w)
    L, R = oracle_lib.lib(), ref_lib.lib()
    for k in range(6):
        m = synth_score_map(k, h, w)
        if k == 3:
            m[:] = -0.75                                         # flat: ties everywhere
        if k == 4:
            m = np.round(m * 2) / 2                              # quantised: many exact ties
        if k == 5:
            m = np.cumsum(m, axis=1).astype(np.float32) * 0.05   # ramps
        w4 = np.array([rng.uniform(0.01, 0.08), rng.uniform(-0.03, 0.03), rng.uniform(0.01, 0.08), rng.uniform(-0.03, 0.03)], np.float32)
        ax, ay = int(rng.integers(-8, 9)), int(rng.integers(-11, 13))
        for dt, ofn, rfn in ((np.float32, L.orc_dt2d_f32, R.ref_dt2d_f32), (np.float64, L.orc_dt2d_f64, R.ref_dt2d_f64)):
            src = np.ascontiguousarray(m, dt)
            o, ox, oy = np.empty((h, w), dt), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
            r, rx, ry = np.empty((h, w), dt), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
            ofn(src.reshape(-1), h, w, w4, ax, ay, 0, o.reshape(-1), ox.reshape(-1), oy.reshape(-1))
            rfn(src.reshape(-1), h, w, w4, ax, ay, r.reshape(-1), rx.reshape(-1), ry.reshape(-1))
            assert np.array_equal(o, r) and np.array_equal(ox, rx) and np.array_equal(oy, ry), (k, dt)


def test_distance_transform_with_non_finite_samples_equals_reference_source():
    """NaN and +-inf samples: whatever the reference's comparisons make of them (nothing is popped past a NaN break point, the scan
    never advances beyond one), the oracle's restatement makes the same -- the pin behind the kernels' literal fallback for such lines."""
    rng = np.random.default_rng(99)
    L, R = oracle_lib.lib(), ref_lib.lib()
    for h, w in ((9, 31), (40, 57)):
        for k in range(4):
            m = (rng.standard_normal((h, w)) * 0.3).astype(np.float32)
            for _ in range(6):
                m[rng.integers(0, h), rng.integers(0, w)] = (np.nan, np.inf, -np.inf)[(k + _) % 3]
            w4 = np.array([rng.uniform(0.01, 0.08), rng.uniform(-0.03, 0.03), rng.uniform(0.01, 0.08), rng.uniform(-0.03, 0.03)], np.float32)
            ax, ay = int(rng.integers(-4, 5)), int(rng.integers(-5, 6))
            o, ox, oy = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
            r, rx, ry = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
            L.orc_dt2d_f32(m.reshape(-1), h, w, w4, ax, ay, 0, o.reshape(-1), ox.reshape(-1), oy.reshape(-1))
            R.ref_dt2d_f32(m.reshape(-1), h, w, w4, ax, ay, r.reshape(-1), rx.reshape(-1), ry.reshape(-1))
            assert np.array_equal(o, r, equal_nan=True) and np.array_equal(ox, rx) and np.array_equal(oy, ry), (h, w, k)


def test_reduce_max_and_pick_index_equal_reference_source():
    rng = np.random.default_rng(5)
    for K in (1, 2, 5, 6):
        h, w = 23, 31
        v = rng.standard_normal((K, h, w)).astype(np.float32)
        v[:, :4] = np.round(v[:, :4])                            # ties: the first maximum wins
        if K > 1:
            v[1, 5] = -np.inf
        pick = rng.integers(0, 1000, (K, h, w)).astype(np.int32)
        mv, mi, pk = np.empty((h, w), np.float32), np.empty((h, w), np.int32), np.empty((h, w), np.int32)
        ref_lib.lib().ref_reduce_max_pick_f32(v.reshape(-1), pick.reshape(-1), K, h, w, mv.reshape(-1), mi.reshape(-1), pk.reshape(-1))
        best, idx = np.full((h, w), -np.inf, np.float32), np.zeros((h, w), np.int32)
        for k in range(K):                                       # the oracle's rule (strict >, init -inf): pbd_oracle.cpp dp_min
            upd = v[k] > best
            best[upd], idx[upd] = v[k][upd], k
        if K == 1:
            best = v[0]
        assert np.array_equal(mv, best) and np.array_equal(mi, idx)
        assert np.array_equal(pk, np.take_along_axis(pick, idx[None], 0)[0])


@pytest.mark.parametrize("name,shape,gray", [("Person_26parts", (120, 160), False), ("Person_26parts", (97, 133), True),
                                             ("Willowcoffee_5parts", (144, 200), False), ("Face_frontal_sparse", (240, 320), False)])
@pytest.mark.parametrize("precision", [32, 64])
def test_hog_pyramid_equals_reference_source(name, shape, gray, precision):
    fm = load_flat(name)
    img = synth_frame(77, *shape)
    if gray:
        img = np.ascontiguousarray(img[:, :, 1])
    ref = ref_lib.hog_pyramid(img, fm.sbin, fm.interval, fm.flen, fm.norient, precision)
    O = oracle_lib.OracleDetector(fm, precision)
    O.run(img, 1, 1)
    assert O.nlevels() == len(ref) > 0
    for l, (feat, scale) in enumerate(ref):
        assert O.level_info(l)["scale"] == scale
        assert np.array_equal(O.features(l), feat), l


def _random_levels(fm, rng, ohow):
    resp = [[(rng.standard_normal(s) * 0.3).astype(np.float32) for _ in range(fm.nfilters())] for s in ohow]
    for l in range(len(ohow)):
        resp[l][0][:] = np.round(resp[l][0] * 4) / 4             # ties in the first filter's map
    return resp


@pytest.mark.parametrize("name", ["Person_26parts", "Face_frontal_sparse", "Person_8parts"])
@pytest.mark.parametrize("precision", [32, 64])
def test_dynamic_program_min_argmin_equal_reference_source(name, precision):
    """Models for which the reference is defined (no part with fewer mixtures than its parent: T4)."""
    fm = load_flat(name)
    rng = np.random.default_rng(len(name) + precision)
    ohow = [(19, 27), (11, 8), (3, 5)]
    scales = np.array([4.0, 5.04, 8.0], np.float32)
    resp = _random_levels(fm, rng, ohow)
    O = oracle_lib.OracleDetector(fm, precision)
    R = ref_lib.RefDP(fm, precision)
    O.set_levels(ohow, scales)
    R.set_levels(ohow, scales)
    for l in range(len(ohow)):
        for f in range(fm.nfilters()):
            O.set_response(l, f, resp[l][f])
            R.set_response(l, f, resp[l][f])
    O.run(None, 3, 3)
    rv = np.sort(np.concatenate([O.rootv(l, c).ravel() for l in range(len(ohow)) for c in range(len(fm.comps))]))
    thr = float(0.5 * (float(rv[-41]) + float(rv[-40])))
    O.set_thresh(thr)
    O.run(None, 4, 4)
    n = R.run(thr)
    for l in range(len(ohow)):
        for c, comp in enumerate(fm.comps):
            v, i = R.root(l, c)
            assert np.array_equal(O.rootv(l, c).astype(np.float64), v) and np.array_equal(O.rooti(l, c), i)
            for p in range(1, len(comp)):
                for pm in range(len(comp[comp[p].parentid].filterid)):
                    for a, b in zip(O.backptr(l, c, p, pm), R.backptr(l, c, p, pm)):
                        assert np.array_equal(a, b), (l, c, p, pm)
    oc = O.candidates()
    assert n == len(oc) and n >= 30
    key = lambda comp, rects, score: (comp, float(score), rects.tobytes())
    got = sorted(key(comp, rects, conf[0]) for comp, rects, conf in R.candidates())
    want = sorted(key(o["component"], o["rects"], o["score"]) for o in oc)
    assert got == want
    for comp, rects, conf in R.candidates():
        assert np.all(conf[1:] == 0.0)                           # only the root carries a confidence, :241-244


def test_candidate_sort_and_nms_equal_reference_source():
    fm = load_flat("Person_26parts")
    rng = np.random.default_rng(3)
    ohow = [(24, 32)]
    scales = np.array([4.0], np.float32)
    resp = _random_levels(fm, rng, ohow)
    R = ref_lib.RefDP(fm, 32)
    O = oracle_lib.OracleDetector(fm, 32)
    R.set_levels(ohow, scales); O.set_levels(ohow, scales)
    for f in range(fm.nfilters()):
        R.set_response(0, f, resp[0][f]); O.set_response(0, f, resp[0][f])
    O.run(None, 3, 3)
    thr = float(np.sort(O.rootv(0).ravel())[-120])
    O.set_thresh(thr); O.run(None, 4, 4)
    assert R.run(thr) == len(O.candidates())
    oc = sorted(O.candidates(), key=lambda o: -float(o["score"]))
    assert len({float(o["score"]) for o in oc}) == len(oc)       # distinct scores: std::sort's order is then defined
    im_h, im_w = 110, 140
    for overlap in (0.0, 0.2, 0.6):
        rects, scores = R.sort_nms(im_h, im_w, overlap)
        flat = np.ascontiguousarray(np.stack([o["rects"] for o in oc]).reshape(-1), np.int32)
        keep = np.zeros(len(oc), np.int32)
        k = oracle_lib.lib().orc_nms(flat, len(oc), 26, im_h, im_w, overlap, keep)
        assert k == len(rects) > 0
        for j in range(k):
            assert np.array_equal(oc[keep[j]]["rects"], rects[j]) and np.float32(oc[keep[j]]["score"]) == scores[j]


@pytest.mark.parametrize("h,w,sz", [(1, 1, 1), (7, 5, 1), (24, 31, 2), (58, 78, 3), (40, 40, 5), (13, 90, 9)])
def test_rootmap_nms_equals_reference_source(h, w, sz):
    """nonMaximaSuppression of src/nms.cpp (unmasked, and masked by a threshold that leaves every block an eligible element)."""
    rng = np.random.default_rng(h * 17 + w + sz)
    L, R = oracle_lib.lib(), ref_lib.lib()
    for k in range(4):
        m = synth_score_map(40 + k, h, w)
        if k == 1:
            m = np.round(m * 2) / 2                              # plateaus: ties are not maxima (strict >)
        if k == 2:
            m[:] = 0.5                                           # constant image: no local maxima
        for mask in (None, (m > np.float32(-10.0)).astype(np.uint8) * 255):
            o, r = np.empty((h, w), np.uint8), np.empty((h, w), np.uint8)
            mp = None if mask is None else mask.ctypes.data
            L.orc_rootmap_nms(m.reshape(-1), h, w, sz, mp, o.reshape(-1))
            R.ref_rootmap_nms(m.reshape(-1), h, w, sz, mp, r.reshape(-1))
            assert np.array_equal(o, r), (k, mask is None)
            if k == 2 and h % (sz + 1) == 0 and w % (sz + 1) == 0 and h > sz + 1 and w > sz + 1:
                assert not o.any()                               # full blocks only (a truncated border block can have an empty window: 0.5 > 0)
            if k == 0 and h * w > 100:
                assert 0 < np.count_nonzero(o) < h * w / ((sz + 1) ** 2) + 1


def test_depth_pruning_equals_reference_source():
    """SearchSpacePruning<float>::filterCandidatesByDepth + Math::median on real candidates of the person model (boxes kept inside the
    depth image: the reference takes depth(box) unclipped)."""
    fm = load_flat("Person_26parts")
    rng = np.random.default_rng(9)
    ohow = [(20, 26)]
    scales = np.array([4.0], np.float32)
    resp = _random_levels(fm, rng, ohow)
    R = ref_lib.RefDP(fm, 32)
    O = oracle_lib.OracleDetector(fm, 32)
    R.set_levels(ohow, scales); O.set_levels(ohow, scales)
    for f in range(fm.nfilters()):
        R.set_response(0, f, resp[0][f]); O.set_response(0, f, resp[0][f])
    O.run(None, 3, 3)
    thr = float(np.sort(O.rootv(0).ravel())[-150])
    assert R.run(thr) > 100
    cands = R.candidates()
    rects = np.stack([c[1] for c in cands]).astype(np.int32)
    im_h, im_w = 400, 400
    inside = np.array([(r[:, 0] >= 0).all() and (r[:, 1] >= 0).all() and ((r[:, 0] + r[:, 2]) <= im_w).all() and ((r[:, 1] + r[:, 3]) <= im_h).all()
                       and (r[:, 2] > 0).all() and (r[:, 3] > 0).all() for r in rects])
    assert inside.sum() > 30
    yy, xx = np.mgrid[0:im_h, 0:im_w]
    depth = (2.0 + 0.0003 * xx + 0.6 * (xx > 85) + 0.004 * rng.standard_normal((im_h, im_w))).astype(np.float32)   # a depth step
    depth[rng.random((im_h, im_w)) < 0.3] = 0.0               # holes: no depth reading
    comp = fm.comps[0]
    parent = np.array([p.parentid for p in comp], np.int32)
    anchor0 = np.array([fm.anchors[p.defid[0]] if p.defid else (0, 0) for p in comp], np.int32).reshape(-1)
    for zf in (0.005, 0.03, 0.2):
        want = R.filter_by_depth(depth, zf)
        keep = np.zeros(len(cands), np.int32)
        oracle_lib.lib().orc_filter_by_depth(np.ascontiguousarray(rects).reshape(-1), len(cands), 26, parent, anchor0, depth.reshape(-1), im_h, im_w, zf, keep)
        assert np.array_equal(keep[inside], want[inside]), zf
    assert 0 < R.filter_by_depth(depth, 0.03)[inside].sum() < inside.sum()
