"""Regenerates the committed fixtures from the reference's model files (run in the build container only;
/root/reference does not exist on the GPU box).

  *.pbdm                 compact binary copies of reference models/*.xml, written by the product's own XML
                         loader (checked field-by-field against cv2.FileStorage in tests/test_model_loader.py)
  Person_26parts_flat.npz  the flat arrays of the north-star model as read by cv2.FileStorage (tests/refmodel.py): lets bench.py's
                         reference arm build the CPU oracle without loading the product library
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


if __name__ == "__main__":
    main()
