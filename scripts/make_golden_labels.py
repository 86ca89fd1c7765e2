"""Pin oracle/labels_oracle.py against the REAL reference: VoxelPostprocessor.generate_label_airv2x /
collate_batch_airv2x imported from /root/reference, with the reference's Cython helper `opencood/utils/box_overlaps.pyx`
compiled into a temp dir (the reference tree is read-only) and injected as `opencood.utils.box_overlaps`. Writes
tests/golden/labels.npz (sparse: positive anchors, their targets / classes, the anchors that are NOT negative).

    python scripts/make_golden_labels.py
"""
import importlib.util
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import labels_oracle as LO, postprocess_oracle as PO, ref_import  # noqa: E402

SEEDS = (201, 202, 203)


def build_box_overlaps():
    tmp = tempfile.mkdtemp(prefix="a2x_box_overlaps_")
    shutil.copy(os.path.join(ref_import.REF_ROOT, "opencood", "utils", "box_overlaps.pyx"), tmp)
    with open(os.path.join(tmp, "setup.py"), "w") as f:
        f.write("from setuptools import setup\nfrom Cython.Build import cythonize\nimport numpy\n"
                "setup(ext_modules=cythonize('box_overlaps.pyx'), include_dirs=[numpy.get_include()])\n")
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, check=True, capture_output=True)
    so = [f for f in os.listdir(tmp) if f.startswith("box_overlaps") and f.endswith(".so")][0]
    spec = importlib.util.spec_from_file_location("opencood.utils.box_overlaps", os.path.join(tmp, so))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref_import.install()
    sys.modules["opencood.utils.box_overlaps"] = build_box_overlaps()
    from opencood.data_utils.post_processor.voxel_postprocessor import VoxelPostprocessor

    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_where2com.json")))
    params = cfg["postprocess"]
    post = VoxelPostprocessor(params, "airv2x", True)
    anchors = post.generate_anchor_box()
    assert np.array_equal(anchors, PO.generate_anchor_box(params["anchor_args"], params["order"]))
    out = {"seeds": np.array(SEEDS)}
    ref_list, ora_list = [], []
    for s in SEEDS:
        box, mask, cls = LO.synth_gt(params, s)
        ref = post.generate_label_airv2x(gt_box_center=box, anchors=anchors, mask=mask, class_ids_padded=cls)
        ora = LO.generate_label(box, mask, cls, anchors, params["target_args"]["pos_threshold"],
                                params["target_args"]["neg_threshold"])
        for k in ("pos_equal_one", "neg_equal_one", "targets", "cls_labels"):
            assert np.array_equal(ref[k], ora[k]), (s, k)
        ref_list.append(ref)
        ora_list.append(ora)
        pos = np.flatnonzero(ref["pos_equal_one"].reshape(-1))
        H, W, A = ref["pos_equal_one"].shape
        print("seed %d: %d gt -> %d positive anchors, %d negative of %d" % (s, int(mask.sum()), pos.size,
                                                                          int(ref["neg_equal_one"].sum()), H * W * A))
        out["pos_idx_%d" % s] = pos
        out["pos_targets_%d" % s] = ref["targets"].reshape(H * W * A, 7)[pos]
        out["pos_cls_%d" % s] = ref["cls_labels"].reshape(-1)[pos]
        out["not_neg_idx_%d" % s] = np.flatnonzero(ref["neg_equal_one"].reshape(-1) == 0)
    a, b = VoxelPostprocessor.collate_batch_airv2x(ref_list), LO.collate(ora_list)
    for k in a:
        assert a[k].dtype == b[k].dtype and bool((a[k] == b[k]).all()), k
    dst = os.path.join(ROOT, "tests", "golden", "labels.npz")
    np.savez_compressed(dst, **out)
    print("oracle == reference on %d frames; wrote %s" % (len(SEEDS), dst))


if __name__ == "__main__":
    main()
