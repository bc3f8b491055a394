"""Regenerates the committed fixtures from the reference's model files (run in the build container only;
/root/reference does not exist on the GPU box).

  *.pbdm                 compact binary copies of reference models/*.xml, written by the product's own XML
                         loader (checked field-by-field against cv2.FileStorage in tests/test_model_loader.py)
  Person_26parts_flat.npz  the flat arrays of the north-star model as read by cv2.FileStorage (tests/refmodel.py): lets bench.py's
                         reference arm build the CPU oracle without loading the product library
  ref_golden.npz         outputs of THE REFERENCE'S OWN SOURCES (oracle/_ref/libpbd_ref.so: HOGFeatures.cpp, DistanceTransform.hpp,
                         DynamicProgram.cpp, nms.cpp compiled unmodified against oracle/ref_shim) on seeded inputs: lets the GPU
                         parity tests compare the CUDA path with reference-source output on the GPU box, where /root/reference is absent
  oracle_golden.npz      outputs of the CPU oracle on small seeded inputs, produced here where the oracle's
                         OpenCV-dependent stages were pinned bit-exactly against cv2 4.13
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from partsbaseddetector_b200 import FileStorageModel  # noqa: E402
from partsbaseddetector_b200.synth import synth_frame  # noqa: E402
import oracle_lib  # noqa: E402
import refmodel  # noqa: E402

MODELS = ["Person_26parts", "Willowcoffee_5parts", "Person_8parts", "Face_frontal_sparse", "Face_99filters"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    for name in MODELS:
        m = FileStorageModel()
        assert m.deserialize(os.path.join(refmodel.REF_MODELS, name + ".xml"))
        m.save_bin(os.path.join(HERE, name + ".pbdm"))
        print("wrote", name + ".pbdm")
    out = {}
    # person model, 160x120 synthetic frame: every stage boundary
    fm = refmodel.load_xml_cv2(os.path.join(refmodel.REF_MODELS, "Person_26parts.xml"))
    np.savez_compressed(os.path.join(HERE, "Person_26parts_flat.npz"), name=np.array(fm.name), thresh=np.float32(fm.thresh), **fm.to_arrays())
    print("wrote Person_26parts_flat.npz")
    D = oracle_lib.OracleDetector(fm, 32)
    img = synth_frame(7, 120, 160)
    D.run(img, 1, 3)
    rv = np.concatenate([D.rootv(l).ravel() for l in range(D.nlevels())])
    thr = float(np.sort(rv)[-40])
    D.set_thresh(thr)
    D.run(None, 4, 4)
    cands = D.candidates()
    out["p26_thresh"] = np.float64(thr)
    out["p26_nlevels"] = np.int32(D.nlevels())
    out["p26_image2"] = D.image(2)
    out["p26_feat0"] = D.features(0)
    out["p26_resp0_f17"] = D.response(0, 17)
    out["p26_rootv0"] = D.rootv(0)
    out["p26_rooti0"] = D.rooti(0)
    ix, iy, ik = D.backptr(0, 0, 3, 2)
    out["p26_ix_p3m2"], out["p26_iy_p3m2"], out["p26_ik_p3m2"] = ix, iy, ik
    out["p26_cand_xyms"] = np.array([[c["level"]] + list(c["x"]) + list(c["y"]) + list(c["m"]) for c in cands], np.int32)
    out["p26_cand_scores"] = np.array([c["score"] for c in cands], np.float32)
    out["p26_cand_rects"] = np.array([c["rects"] for c in cands], np.int32)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote oracle_golden.npz", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
    write_ref_golden(fm)


def write_ref_golden(fm):
    """Vectors produced by the reference's compiled sources (tests/ref_lib.py)."""
    import ref_lib
    from partsbaseddetector_b200.synth import synth_score_map
    assert ref_lib.available()
    g = {}
    # HOGFeatures<float>::pyramid on a 120x160 frame: every level's features and scale
    img = synth_frame(11, 120, 160)
    pyr = ref_lib.hog_pyramid(img, fm.sbin, fm.interval, fm.flen, fm.norient, 32)
    g["hog_nlevels"] = np.int32(len(pyr))
    for l, (feat, scale) in enumerate(pyr):
        g["hog_feat%d" % l] = feat
    g["hog_scales"] = np.array([s for _, s in pyr], np.float32)
    # DistanceTransform<float>::compute on 4 maps (incl. a flat and a quantised one)
    rng = np.random.default_rng(77)
    maps = np.stack([synth_score_map(30 + i, 61, 83) for i in range(4)])
    maps[2] = 0.5
    maps[3] = np.round(maps[3] * 2) / 2
    defw = np.stack([rng.uniform(0.01, 0.06, 4), rng.uniform(-0.02, 0.02, 4), rng.uniform(0.01, 0.06, 4), rng.uniform(-0.02, 0.02, 4)], axis=1).astype(np.float32)
    anch = np.stack([rng.integers(-5, 6, 4), rng.integers(-4, 7, 4)], axis=1).astype(np.int32)
    outs = np.empty_like(maps); ixs = np.empty(maps.shape, np.int32); iys = np.empty(maps.shape, np.int32)
    for i in range(4):
        ref_lib.lib().ref_dt2d_f32(maps[i].reshape(-1), 61, 83, defw[i], int(anch[i, 0]), int(anch[i, 1]), outs[i].reshape(-1), ixs[i].reshape(-1), iys[i].reshape(-1))
    g["dt_in"], g["dt_defw"], g["dt_anchor"], g["dt_out"], g["dt_ix"], g["dt_iy"] = maps, defw, anch, outs, ixs, iys
    # DynamicProgram<float>::min + argmin of the person model on seeded responses (2 levels), root-map NMS of level 0
    ohow = [(22, 30), (9, 13)]
    scales = np.array([4.0, 8.0], np.float32)
    R = ref_lib.RefDP(fm, 32)
    R.set_levels(ohow, scales)
    seeds = np.arange(fm.nfilters())
    for l, shp in enumerate(ohow):
        for f in range(fm.nfilters()):
            R.set_response(l, f, (np.random.default_rng(1000 * l + f).standard_normal(shp) * 0.3).astype(np.float32))
    probe = R.run(1e9)
    assert probe == 0
    rv = np.sort(np.concatenate([R.root(l)[0].ravel() for l in range(2)]))
    thr = float(0.5 * (rv[-61] + rv[-60]))
    n = R.run(thr)
    g["dp_thresh"], g["dp_ohow"], g["dp_scales"] = np.float64(thr), np.array(ohow, np.int32), scales
    for l in range(2):
        v, i = R.root(l)
        g["dp_rootv%d" % l], g["dp_rooti%d" % l] = v.astype(np.float32), i
    for (p, pm) in ((1, 0), (3, 2), (14, 4), (25, 1)):
        ix, iy, ik = R.backptr(0, 0, p, pm)
        g["dp_bp_p%d_m%d" % (p, pm)] = np.stack([ix, iy, ik])
    cands = R.candidates()
    order = sorted(range(n), key=lambda i: (float(cands[i][2][0]), cands[i][1].tobytes()))
    g["dp_cand_rects"] = np.stack([cands[i][1] for i in order]).astype(np.int32)
    g["dp_cand_scores"] = np.array([cands[i][2][0] for i in order], np.float32)
    keep = np.empty(ohow[0], np.uint8)
    ref_lib.lib().ref_rootmap_nms(g["dp_rootv0"].reshape(-1), ohow[0][0], ohow[0][1], 2, None, keep.reshape(-1))
    g["dp_rootnms2_level0"] = keep
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **g)
    print("wrote ref_golden.npz (%d arrays from the reference's compiled sources, %d candidates)" % (len(g), n))


if __name__ == "__main__":
    main()
