"""Pins the CPU oracle (oracle/pbd_oracle.cpp) -- the reference ships no golden vectors for this path
(SURVEY.md section 4), so the restatement is checked against the authorities available here:
cv2 4.13 for the OpenCV routines the reference calls, brute force for the distance transform and the DP,
an independent numpy HOG, and the committed golden vectors generated in the build container."""
import itertools
import os

import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN, load_flat
from partsbaseddetector_b200.flatmodel import FlatModel, FlatPart
from partsbaseddetector_b200.synth import synth_frame

cv2 = pytest.importorskip("cv2")
L = oracle_lib.lib()


# ---------------------------------------------------------------- cv::resize / cv::pyrDown (bit-exact)
@pytest.mark.parametrize("shape,dst", [((480, 640), (381, 508)), ((480, 640), (302, 403)), ((240, 320), (240, 320)),
                                       ((97, 131), (61, 83)), ((1080, 1920), (857, 1524)), ((33, 47), (26, 37))])
def test_resize_bitexact_vs_cv2(shape, dst):
    rng = np.random.default_rng(shape[0] * 7 + dst[1])
    for img in (rng.integers(0, 256, shape + (3,)).astype(np.uint8), synth_frame(3, *shape)):
        out = np.empty(dst + (3,), np.uint8)
        L.orc_resize_u8(img.reshape(-1), shape[0], shape[1], 3, out.reshape(-1), dst[0], dst[1])
        assert np.array_equal(out, cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_LINEAR))


def test_resize_gray_bitexact_vs_cv2():
    img = np.random.default_rng(5).integers(0, 256, (120, 160)).astype(np.uint8)
    out = np.empty((95, 127), np.uint8)
    L.orc_resize_u8(img.reshape(-1), 120, 160, 1, out.reshape(-1), 95, 127)
    assert np.array_equal(out, cv2.resize(img, (127, 95), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("shape", [(480, 640), (381, 508), (5, 7), (3, 3), (151, 202), (96, 127)])
def test_pyrdown_bitexact_vs_cv2(shape):
    img = np.random.default_rng(shape[0]).integers(0, 256, shape + (3,)).astype(np.uint8)
    out = np.empty(((shape[0] + 1) // 2, (shape[1] + 1) // 2, 3), np.uint8)
    L.orc_pyrdown_u8(img.reshape(-1), shape[0], shape[1], 3, out.reshape(-1))
    assert np.array_equal(out, cv2.pyrDown(img))


def test_pyramid_geometry_tables():
    # SURVEY.md appendix C level tables
    def levels(h, w, sbin, interval):
        wh = np.zeros(2 * 96, np.int32)
        sc = np.zeros(96, np.float32)
        n = L.orc_pyramid_geometry(h, w, sbin, interval, 96, wh, sc)
        return n, wh[:2 * n].reshape(n, 2), sc[:n]
    n, wh, sc = levels(480, 640, 4, 3)
    assert n == 14
    assert wh.tolist()[:7] == [[640, 480], [508, 381], [403, 302], [320, 240], [254, 191], [202, 151], [160, 120]]
    assert wh.tolist()[-1] == [32, 24]
    assert sc[0] == 4.0 and sc[3] == 8.0 and sc[6] == 16.0
    assert levels(1080, 1920, 4, 3)[0] == 18
    assert levels(240, 320, 4, 10)[0] == 36


# ---------------------------------------------------------------- responses vs cv2.filter2D
@pytest.mark.parametrize("kh,kw", [(5, 5), (6, 6), (4, 4), (7, 11)])
def test_convolve_vs_filter2d(kh, kw):
    rng = np.random.default_rng(kh * 10 + kw)
    oh, ow, C = 23, 31, 32
    feat = rng.random((oh, ow, C)).astype(np.float32)
    feat[:, :, 31] = 0
    filt = (rng.standard_normal((kh, kw, C)) * 0.01).astype(np.float32)
    out = np.empty((oh, ow), np.float32)
    L.orc_convolve_f32(feat.reshape(-1), oh, ow, C, filt.reshape(-1), kh, kw, out.reshape(-1))
    ref = np.zeros((oh, ow), np.float64)
    ay, ax = kh // 2, kw // 2
    for c in range(C):
        border = 1.0 if c == C - 1 else 0.0
        padded = cv2.copyMakeBorder(feat[:, :, c], ay, kh - 1 - ay, ax, kw - 1 - ax, cv2.BORDER_CONSTANT, value=border)
        r = cv2.filter2D(padded.astype(np.float64), cv2.CV_64F, filt[:, :, c].astype(np.float64), anchor=(ax, ay),
                         borderType=cv2.BORDER_CONSTANT)
        ref += r[ay:ay + oh, ax:ax + ow]
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)
    # double-precision oracle agrees to 1e-12
    out64 = np.empty((oh, ow), np.float64)
    L.orc_convolve_f64(feat.astype(np.float64).reshape(-1), oh, ow, C, filt.astype(np.float64).reshape(-1), kh, kw, out64.reshape(-1))
    assert np.allclose(out64, ref, rtol=1e-11, atol=1e-12)


# ---------------------------------------------------------------- distance transform vs brute force
def brute_dt1d(src, a, b, os_):
    N = len(src)
    q = np.arange(N)[:, None] + os_
    v = np.arange(N)[None, :]
    d = (q - v).astype(np.float64)
    val = src[None, :].astype(np.float64) + a * d * d + b * d
    return val.max(axis=1), val.argmax(axis=1), val


@pytest.mark.parametrize("seed", range(8))
def test_dt1d_vs_bruteforce(seed):
    rng = np.random.default_rng(seed)
    for _ in range(25):
        N = int(rng.integers(1, 90))
        src = rng.standard_normal(N).astype(np.float32)
        a, b = -float(rng.uniform(0.01, 0.08)), float(rng.uniform(-0.02, 0.02))
        os_ = int(rng.integers(-9, 10))
        dst = np.empty(N, np.float32)
        ptr = np.empty(N, np.int32)
        L.orc_dt1d_f32(src, N, a, b, os_, dst, ptr)
        bv, bi, val = brute_dt1d(src, a, b, os_)
        assert np.allclose(dst, bv, rtol=1e-6, atol=1e-6)
        chosen = val[np.arange(N), ptr]
        assert np.all(np.abs(chosen - bv) <= 1e-6 * (1 + np.abs(bv)))       # the chosen parabola attains the max
        second = np.sort(val, axis=1)[:, -2] if N > 1 else bv - 1
        clear = (bv - second) > 1e-5
        assert np.array_equal(ptr[clear], bi[clear])                          # argmax identical away from ties


def test_dt2d_value_and_pointer_composition():
    rng = np.random.default_rng(11)
    M, N = 12, 15
    src = rng.standard_normal((M, N)).astype(np.float32)
    w = np.array([0.013, 0.004, 0.017, -0.009], np.float32)
    osx, osy = 2, -1
    out, ix, iy = (np.empty((M, N), t) for t in (np.float32, np.int32, np.int32))
    L.orc_dt2d_f32(src.reshape(-1), M, N, w, osx, osy, 0, out.reshape(-1), ix.reshape(-1), iy.reshape(-1))
    a0, b0, a1, b1 = -float(w[0]), -float(w[1]), -float(w[2]), -float(w[3])
    yy, xx = np.mgrid[0:M, 0:N]
    best = np.empty((M, N))
    for y in range(M):
        for x in range(N):
            dx, dy = (x + osx) - xx, (y + osy) - yy
            best[y, x] = (src + a0 * dx * dx + b0 * dx + a1 * dy * dy + b1 * dy).max()
    assert np.allclose(out, best, rtol=1e-5, atol=1e-5)
    # reference composition rule (T1): Iy = Iyraw[y][Ix], Ix = row-pass argmax
    tmp, ixr = np.empty((M, N), np.float32), np.empty((M, N), np.int32)
    for y in range(M):
        L.orc_dt1d_f32(src[y].copy(), N, a0, b0, osx, tmp[y], ixr[y])
    iyraw = np.empty((M, N), np.int32)
    for x in range(N):
        d, p = np.empty(M, np.float32), np.empty(M, np.int32)
        L.orc_dt1d_f32(np.ascontiguousarray(tmp[:, x]), M, a1, b1, osy, d, p)
        iyraw[:, x] = p
    assert np.array_equal(ix, ixr)
    assert np.array_equal(iy, np.take_along_axis(iyraw, ixr, axis=1))
    # exact mode: Ix = Ixraw[Iy][x]
    out1, ix1, iy1 = (np.empty((M, N), t) for t in (np.float32, np.int32, np.int32))
    L.orc_dt2d_f32(src.reshape(-1), M, N, w, osx, osy, 1, out1.reshape(-1), ix1.reshape(-1), iy1.reshape(-1))
    assert np.array_equal(out1, out) and np.array_equal(iy1, iyraw)
    assert np.array_equal(ix1, ixr[iyraw, np.arange(N)[None, :]])
    att = src[iy1, ix1] + a0 * ((xx + osx) - ix1) ** 2 + b0 * ((xx + osx) - ix1) + a1 * ((yy + osy) - iy1) ** 2 + b1 * ((yy + osy) - iy1)
    assert np.allclose(att, best, rtol=1e-5, atol=1e-5)                        # exact pointers attain the 2-D max


# ---------------------------------------------------------------- DP vs exhaustive enumeration on a toy model
def toy_model(rng):
    flen = 32
    nf = 6                                       # 3 parts x 2 mixtures: root(0,1) <- part1(2,3) <- part2(4,5)
    filters = [(rng.standard_normal((3, 3 * flen)) * 0.01) for _ in range(nf)]
    biasw = (rng.standard_normal(1 + 4 + 4) * 0.1).astype(np.float32)
    anchors = np.array([[1, -1], [0, 2], [-1, 0], [2, 1]], np.int32)
    defs = np.stack([rng.uniform(0.01, 0.05, 4), rng.uniform(-0.02, 0.02, 4), rng.uniform(0.01, 0.05, 4),
                     rng.uniform(-0.02, 0.02, 4)], axis=1).astype(np.float32)
    comps = [[FlatPart(-1, [0, 1], [0], [0]), FlatPart(0, [2, 3], [1, 3], [0, 1]), FlatPart(1, [4, 5], [5, 7], [2, 3])]]
    return FlatModel("toy", 3, -10.0, 4, 18, flen, filters, biasw, anchors, defs, comps)


def test_dp_root_score_vs_exhaustive():
    rng = np.random.default_rng(3)
    fm = toy_model(rng)
    oh, ow = 5, 6
    D = oracle_lib.OracleDetector(fm, 64)
    D.set_levels([[oh, ow]], [4.0])
    resp = [rng.standard_normal((oh, ow)) for _ in range(6)]
    for f in range(6):
        D.set_response(0, f, resp[f])
    D.set_backptr_mode(1)
    D.run(None, 3, 4)
    rootv, rooti = D.rootv(0), D.rooti(0)
    cells = list(itertools.product(range(oh), range(ow)))

    def pen(did, px, py, cx, cy):
        w, (ax, ay) = fm.defs[did].astype(np.float64), fm.anchors[did]
        dx, dy = (px + ax) - cx, (py + ay) - cy
        return -w[0] * dx * dx - w[1] * dx - w[2] * dy * dy - w[3] * dy

    def child_msg(part, pm, px, py):                       # max over child mixture and location (recursive)
        P = fm.comps[0][part]
        best = -np.inf
        for mm, fid in enumerate(P.filterid):
            for (cy, cx) in cells:
                s = resp[fid][cy, cx] + float(fm.biasw[P.biasid[mm] + pm]) + pen(P.defid[mm], px, py, cx, cy)
                for ch in range(len(fm.comps[0])):
                    if fm.comps[0][ch].parentid == part:
                        s += child_msg(ch, mm, cx, cy)
                best = max(best, s)
        return best

    for (y, x) in cells[::3]:
        vals = [resp[fid][y, x] + float(fm.biasw[0]) + child_msg(1, m, x, y) for m, fid in enumerate(fm.comps[0][0].filterid)]
        assert abs(max(vals) - rootv[y, x]) < 1e-9
        assert int(np.argmax(vals)) == rooti[y, x]


# ---------------------------------------------------------------- HOG vs an independent numpy formulation
def numpy_hog(img, sbin=4):
    """Vectorised float64 HOG after Felzenszwalb's features.cc as used by the reference (HWC, BGR, 32 dims)."""
    im = img.astype(np.float64)
    rows, cols = im.shape[:2]
    bw, bh = int(np.floor(cols / sbin + 0.5)), int(np.floor(rows / sbin + 0.5))
    vis_w, vis_h = bw * sbin, bh * sbin
    ys, xs = np.arange(1, vis_h - 1), np.arange(1, vis_w - 1)
    sy, sx = np.minimum(ys, rows - 2), np.minimum(xs, cols - 2)
    dy = im[sy + 1][:, sx] - im[sy - 1][:, sx]
    dx = im[sy][:, sx + 1] - im[sy][:, sx - 1]
    v = dx * dx + dy * dy                                   # (H, W, 3) in B, G, R order
    pick = np.full(v.shape[:2], 2)                          # R wins ties, then G, then B
    pick = np.where(v[:, :, 1] > v[:, :, 2], 1, pick)
    vbest = np.maximum(v[:, :, 2], v[:, :, 1])
    pick = np.where(v[:, :, 0] > vbest, 0, pick)
    ii, jj = np.indices(pick.shape)
    dxs, dys, vs = dx[ii, jj, pick], dy[ii, jj, pick], v[ii, jj, pick]
    uu = np.array([1.000, 0.9397, 0.7660, 0.5000, 0.1736, -0.1736, -0.5000, -0.7660, -0.9397])
    vv = np.array([0.000, 0.3420, 0.6428, 0.8660, 0.9848, 0.9848, 0.8660, 0.6428, 0.3420])
    dots = np.stack([u * dxs + w * dys for u, w in zip(uu, vv)], axis=-1)
    allv = np.concatenate([dots, -dots], axis=-1)           # orientation o and o+9
    # reference scan: o ascending, +dot tested before -dot, strict >
    order = np.array([k for o in range(9) for k in (o, o + 9)])
    best = np.zeros(pick.shape)
    bo = np.zeros(pick.shape, int)
    for k in order:
        better = allv[:, :, k] > best
        best = np.where(better, allv[:, :, k], best)
        bo = np.where(better, k, bo)
    mag = np.sqrt(vs)
    hist = np.zeros((bh, bw, 18))
    yp, xp = (ys + 0.5) / sbin - 0.5, (xs + 0.5) / sbin - 0.5
    iyp, ixp = np.floor(yp).astype(int), np.floor(xp).astype(int)
    vy0, vx0 = yp - iyp, xp - ixp
    for (oy, wy) in ((0, 1 - vy0), (1, vy0)):
        for (ox, wx) in ((0, 1 - vx0), (1, vx0)):
            by, bx = (iyp + oy)[:, None] + 0 * ixp[None, :], (ixp + ox)[None, :] + 0 * iyp[:, None]
            ok = (by >= 0) & (by < bh) & (bx >= 0) & (bx < bw)
            np.add.at(hist, (by[ok], bx[ok], bo[ok]), (wy[:, None] * wx[None, :] * mag)[ok])
    norm = ((hist[:, :, :9] + hist[:, :, 9:]) ** 2).sum(-1)
    oh, ow = max(bh - 2, 0), max(bw - 2, 0)
    feat = np.zeros((oh, ow, 32))
    eps = 0.0001
    ns = []
    for (ay, ax) in ((1, 1), (0, 1), (1, 0), (0, 0)):
        s = norm[ay:ay + oh, ax:ax + ow] + norm[ay:ay + oh, ax + 1:ax + 1 + ow] + norm[ay + 1:ay + 1 + oh, ax:ax + ow] + norm[ay + 1:ay + 1 + oh, ax + 1:ax + 1 + ow]
        ns.append(1.0 / np.sqrt(s + eps))
    h = hist[1:1 + oh, 1:1 + ow]
    hs = [np.minimum(h * n[:, :, None], 0.2) for n in ns]
    feat[:, :, :18] = 0.5 * sum(hs)
    hsum = h[:, :, :9] + h[:, :, 9:]
    feat[:, :, 18:27] = 0.5 * sum(np.minimum(hsum * n[:, :, None], 0.2) for n in ns)
    for k in range(4):
        feat[:, :, 27 + k] = 0.2357 * hs[k].sum(-1)
    return feat


@pytest.mark.parametrize("shape", [(120, 160), (97, 131), (64, 64)])
def test_hog_vs_numpy(shape):
    img = synth_frame(shape[0], *shape)
    oh, ow = oracle_lib.C.c_int(), oracle_lib.C.c_int()
    L.orc_hog_dims(shape[0], shape[1], 4, oracle_lib.C.byref(oh), oracle_lib.C.byref(ow))
    ref = numpy_hog(img)
    assert ref.shape[:2] == (oh.value, ow.value)
    f64 = np.empty(ref.size, np.float64)
    L.orc_hog_f64(img.reshape(-1), shape[0], shape[1], 3, 4, 18, 32, f64)
    f64 = f64.reshape(ref.shape)
    # the reference rounds n_k to T and takes 1.0f/sqrt: agreement to ~1e-12 in double, 1e-5 in float
    assert np.allclose(f64, ref, rtol=1e-9, atol=1e-12)
    f32 = np.empty(ref.size, np.float32)
    L.orc_hog_f32(img.reshape(-1), shape[0], shape[1], 3, 4, 18, 32, f32)
    f32 = f32.reshape(ref.shape)
    assert np.allclose(f32, ref, rtol=2e-5, atol=2e-6)
    assert np.all(f32[:, :, 31] == 0) and f32[:, :, :27].max() <= 0.4 + 1e-6 and f32[:, :, 27:31].max() <= 0.8485 + 1e-4


# ---------------------------------------------------------------- committed golden vectors
def test_oracle_reproduces_committed_golden():
    g = np.load(os.path.join(GOLDEN, "oracle_golden.npz"))
    fm = load_flat("Person_26parts")
    D = oracle_lib.OracleDetector(fm, 32)
    D.run(synth_frame(7, 120, 160), 1, 3)
    D.set_thresh(float(g["p26_thresh"]))
    D.run(None, 4, 4)
    assert D.nlevels() == int(g["p26_nlevels"])
    assert np.array_equal(D.image(2), g["p26_image2"])
    assert np.array_equal(D.features(0), g["p26_feat0"])
    assert np.array_equal(D.response(0, 17), g["p26_resp0_f17"])
    assert np.array_equal(D.rootv(0), g["p26_rootv0"]) and np.array_equal(D.rooti(0), g["p26_rooti0"])
    ix, iy, ik = D.backptr(0, 0, 3, 2)
    assert np.array_equal(ix, g["p26_ix_p3m2"]) and np.array_equal(iy, g["p26_iy_p3m2"]) and np.array_equal(ik, g["p26_ik_p3m2"])
    c = D.candidates()
    assert np.array_equal(np.array([[k["level"]] + list(k["x"]) + list(k["y"]) + list(k["m"]) for k in c], np.int32), g["p26_cand_xyms"])
    assert np.array_equal(np.array([k["score"] for k in c], np.float32), g["p26_cand_scores"])
    assert np.array_equal(np.array([k["rects"] for k in c], np.int32), g["p26_cand_rects"])


def test_oracle_thread_count_invariance():
    fm = load_flat("Willowcoffee_5parts")
    img = synth_frame(2, 96, 128)
    res = []
    for nt in ("1", "3"):
        os.environ["OMP_NUM_THREADS"] = nt      # read at first parallel region of a new team only; use omp via env for the subprocess
        import subprocess, sys, json
        code = ("import sys,numpy as np;sys.path.insert(0,'tests');sys.path.insert(0,'.');import oracle_lib;from conftest import load_flat;"
                "from partsbaseddetector_b200.synth import synth_frame;D=oracle_lib.OracleDetector(load_flat('Willowcoffee_5parts'),32);"
                "D.run(synth_frame(2,96,128),1,3);import hashlib;"
                "print(hashlib.sha256(b''.join(D.rootv(l).tobytes() for l in range(D.nlevels()))).hexdigest())")
        res.append(subprocess.check_output([sys.executable, "-c", code], cwd=os.path.dirname(GOLDEN) + "/..", env=dict(os.environ)).strip())
    assert res[0] == res[1]


def test_oracle_equals_vectors_from_the_reference_sources():
    """tests/golden/ref_golden.npz was produced by the reference's own compiled sources (tests/golden/make_golden.py through
    oracle/_ref): the oracle reproduces it bit for bit -- also where /root/reference and oracle/_ref do not exist."""
    import os
    from conftest import GOLDEN, load_flat
    g = np.load(os.path.join(GOLDEN, "ref_golden.npz"))
    fm = load_flat("Person_26parts")
    O = oracle_lib.OracleDetector(fm, 32)
    O.run(synth_frame(11, 120, 160), 1, 1)
    assert O.nlevels() == int(g["hog_nlevels"])
    for l in range(O.nlevels()):
        assert np.array_equal(O.features(l), g["hog_feat%d" % l]) and O.level_info(l)["scale"] == g["hog_scales"][l]
    L = oracle_lib.lib()
    for i in range(4):
        m = np.ascontiguousarray(g["dt_in"][i])
        o, x, y = np.empty_like(m), np.empty(m.shape, np.int32), np.empty(m.shape, np.int32)
        L.orc_dt2d_f32(m.reshape(-1), m.shape[0], m.shape[1], np.ascontiguousarray(g["dt_defw"][i]), int(g["dt_anchor"][i, 0]), int(g["dt_anchor"][i, 1]), 0,
                       o.reshape(-1), x.reshape(-1), y.reshape(-1))
        assert np.array_equal(o, g["dt_out"][i]) and np.array_equal(x, g["dt_ix"][i]) and np.array_equal(y, g["dt_iy"][i])
    ohow = [tuple(int(v) for v in r) for r in g["dp_ohow"]]
    O.set_levels(ohow, g["dp_scales"])
    for l, shp in enumerate(ohow):
        for f in range(fm.nfilters()):
            O.set_response(l, f, (np.random.default_rng(1000 * l + f).standard_normal(shp) * 0.3).astype(np.float32))
    O.set_thresh(float(g["dp_thresh"]))
    O.run(None, 3, 4)
    for l in range(2):
        assert np.array_equal(O.rootv(l), g["dp_rootv%d" % l]) and np.array_equal(O.rooti(l), g["dp_rooti%d" % l])
    for (p, pm) in ((1, 0), (3, 2), (14, 4), (25, 1)):
        assert np.array_equal(np.stack(O.backptr(0, 0, p, pm)), g["dp_bp_p%d_m%d" % (p, pm)])
    oc = sorted(O.candidates(), key=lambda o: (float(o["score"]), o["rects"].astype(np.int32).tobytes()))
    assert np.array_equal(np.stack([o["rects"] for o in oc]), g["dp_cand_rects"])
    assert np.array_equal(np.array([o["score"] for o in oc], np.float32), g["dp_cand_scores"])
    keep = np.empty(ohow[0], np.uint8)
    L.orc_rootmap_nms(np.ascontiguousarray(g["dp_rootv0"]).reshape(-1), ohow[0][0], ohow[0][1], 2, None, keep.reshape(-1))
    assert np.array_equal(keep, g["dp_rootnms2_level0"])
