"""Pin the dataset side (`airv2x-perception_b200/intermediate_fusion_dataset.py`) against the REAL reference class:
`IntermediateFusionDatasetAirv2x.__getitem__` + `collate_batch_train` imported from /root/reference are run on seeded
synthetic `retrieve_base_data` records (tests/dataset_common.py). The reference object is created without its directory
scan (`__new__` + the attributes `__init__` would set from the yaml), `retrieve_base_data` returns the synthetic record and
the spconv voxeliser (absent here) is replaced by the sequential restatement `oracle/voxelize.py` behind the reference's
own `SpVoxelPreprocessor.collate_batch`. Everything else — range pruning, nearest-k selection, object projection, id
de-duplication, label maps, pairwise matrices, priors, camera tensors, the cloud filters — is the reference's code.

Asserts that this repo's dataset emits the same batch (bookkeeping equal, boxes to 1e-9, the raw clouds pushed through the
restated filters + voxeliser equal to the reference's voxel tensors bit for bit) and writes tests/golden/dataset.npz for
the tests that run where /root/reference is absent.

    python scripts/make_golden_dataset.py
"""
import copy
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import a2x_import  # noqa: E402
from oracle import ref_import, voxelize as VX  # noqa: E402

YAML = "airv2x/lidar/det/airv2x_intermediate_where2com.yaml"
CASES = {  # name -> (train, list of synth_scene kwargs)
    "train_b3": (True, [dict(seed=11), dict(seed=12, n_veh=1, n_rsu=0, n_drone=1, far=False),
                        dict(seed=13, n_veh=7, n_rsu=6, n_drone=1, n_obj=60)]),
    "test_b1": (False, [dict(seed=21, n_veh=2, n_rsu=1, n_drone=0)]),
}


def reference_env():
    """stubs for the packages / extensions the dataset module imports but this container lacks; returns the module"""
    import make_golden_labels as MGL
    ref_import.install()
    if "more_itertools" not in sys.modules or isinstance(sys.modules["more_itertools"], MagicMock):
        mi = types.ModuleType("more_itertools")

        def unique_everseen(it):            # the one function used: order-preserving de-duplication
            seen = set()
            for x in it:
                if x not in seen:
                    seen.add(x)
                    yield x
        mi.unique_everseen = unique_everseen
        sys.modules["more_itertools"] = mi
    sys.modules.setdefault("opencood.pcdet_utils.roiaware_pool3d.roiaware_pool3d_cuda", MagicMock())
    if isinstance(sys.modules.get("opencood.utils.box_overlaps"), (type(None), MagicMock)):
        sys.modules["opencood.utils.box_overlaps"] = MGL.build_box_overlaps()
    from opencood.data_utils.datasets.airv2x import intermediate_fusion_dataset as IFD
    return IFD


class _OraclePreprocessor:
    """`SpVoxelPreprocessor` with the absent spconv generator replaced by the sequential restatement"""

    def __init__(self, params, train):
        self.range, self.vs = params["cav_lidar_range"], params["args"]["voxel_size"]
        self.mp = params["args"]["max_points_per_voxel"]
        self.mv = params["args"]["max_voxel_train" if train else "max_voxel_test"]

    def preprocess(self, pcd_np):
        return VX.voxelize(pcd_np, self.range, self.vs, self.mp, self.mv)

    def collate_batch(self, batch):
        from opencood.data_utils.pre_processor.sp_voxel_preprocessor import SpVoxelPreprocessor
        return SpVoxelPreprocessor.collate_batch(self, batch)


def reference_dataset(IFD, hypes, train):
    """the reference class without `BaseDataset.__init__`'s directory scan: the attributes it derives from the yaml"""
    from opencood.data_utils import post_processor
    ds = IFD.IntermediateFusionDatasetAirv2x.__new__(IFD.IntermediateFusionDatasetAirv2x)
    ds.params, ds.visualize, ds.train, ds.training = hypes, False, train, train
    ds.proj_first = hypes["fusion"]["args"]["proj_first"]
    mc = hypes["train_params"]["max_cav"]
    ds.max_cav_veh, ds.max_cav_rsu, ds.max_cav_drone = mc["vehicle"], mc["rsu"], mc["drone"]
    ds.collaborators, ds.active_sensors = hypes["collaborators"], hypes["active_sensors"]
    ds.max_cav_num = sum(mc[c] for c in ds.collaborators)
    fa = hypes["fusion"]["args"]
    ds.veh_data_aug_conf, ds.rsu_data_aug_conf, ds.drone_data_aug_conf = (fa["veh_data_aug_conf"], fa["rsu_data_aug_conf"],
                                                                          fa["drone_data_aug_conf"])
    ds.cur_ego_pose_flag = True
    ds.ego_type = hypes.get("ego_type", "vehicle")
    ds.agent_order = {"vehicle": ["vehicle", "rsu", "drone"], "rsu": ["rsu", "vehicle", "drone"],
                      "drone": ["drone", "vehicle", "rsu"]}[ds.ego_type]
    ds.pre_processor = _OraclePreprocessor(hypes["preprocess"], train)
    ds.post_processor = post_processor.build_postprocessor(hypes["postprocess"], dataset="airv2x", train=train)
    return ds


def run_reference(ds, scenes, seed):
    """(per-scene items, collated batch) of the real class; numpy's global generator seeded like the run under test"""
    np.random.seed(seed)
    items = []
    for i, base in enumerate(scenes):
        ds.retrieve_base_data = lambda idx, cur_ego_pos_flag=True, _b=base, _i=i: (copy.deepcopy(_b), 0, _i)
        items.append(ds[i])
    return items, ds.collate_batch_train(items)


def run_ours(DS, hypes, train, scenes, seed, **kw):
    np.random.seed(seed)
    ds = DS.IntermediateFusionDatasetAirv2x(hypes, False, train, source=[copy.deepcopy(b) for b in scenes], **kw)
    items = [ds[i] for i in range(len(scenes))]
    return ds, items, ds.collate_batch_train(items)


def voxelise_like_the_reference(ours, hypes, train):
    """our raw_points pushed through the restated cloud filters + sequential voxeliser, grouped per agent type the way
    `collate_batch_train` groups them -> {type: {"voxel_features", "voxel_coords", "voxel_num_points"}}"""
    raw = ours["raw_points"]
    pre = _OraclePreprocessor(hypes["preprocess"], train)
    pts, offs, tf = raw["points"].numpy(), raw["offsets"].numpy(), raw["transforms"].numpy()
    out, row = {}, 0
    per_type = {t: [] for t in ("vehicle", "rsu", "drone")}
    for b in range(len(ours["record_len"])):
        for t in ("vehicle", "rsu", "drone"):
            for _ in range(int(ours[t]["record_len"][b])):
                cloud = VX.dataset_points(pts[offs[row]:offs[row + 1]], tf[row], pre.range)
                per_type[t].append(pre.preprocess(cloud))
                row += 1
    assert row == len(offs) - 1
    for t, lst in per_type.items():
        out[t] = VX.collate(lst) if lst else None
    return out


def compare(ref_batch, ours_batch, hypes, train):
    """assert the two collated batches describe the same scenes; returns the max box deviation"""
    r, o = ref_batch["ego"], ours_batch["ego"]
    for k in ("record_len", "pairwise_t_matrix_collab", "img_pairwise_t_matrix_collab", "prior_encoding",
              "spatial_correction_matrix", "object_bbx_mask"):
        assert r[k].dtype == o[k].dtype and r[k].shape == o[k].shape, (k, r[k].dtype, o[k].dtype, r[k].shape, o[k].shape)
    for k in ("record_len", "prior_encoding", "object_bbx_mask"):
        assert torch.equal(r[k], o[k]), k
    for k in ("pairwise_t_matrix_collab", "img_pairwise_t_matrix_collab", "spatial_correction_matrix"):
        assert torch.allclose(r[k], o[k], rtol=0, atol=1e-6 if r[k].dtype == torch.float32 else 1e-12), k
    assert r["object_ids"] == o["object_ids"] and r["class_ids"] == o["class_ids"]
    dev = float((r["object_bbx_center"] - o["object_bbx_center"]).abs().max())
    assert dev < 1e-9, dev
    for k in ("scenario_index_list", "timestamp_key_list", "metadata_path_list", "ego_lidar_pose_list"):
        assert r[k] == o[k], k
    vox = voxelise_like_the_reference(o, hypes, train)
    for t in ("vehicle", "rsu", "drone"):
        assert torch.equal(r[t]["record_len"], o[t]["record_len"]) and list(r[t]["batch_idxs"]) == list(o[t]["batch_idxs"]), t
        rv = r[t]["batch_merged_lidar_features_torch"]
        assert (rv is None) == (vox[t] is None), t
        if rv is not None:
            for k in ("voxel_features", "voxel_coords", "voxel_num_points"):
                assert np.array_equal(rv[k].numpy(), vox[t][k]), (t, k)
        rc, oc = r[t]["batch_merged_cam_inputs"], o[t]["batch_merged_cam_inputs"]
        assert set(rc.keys()) == set(oc.keys()), (t, rc.keys(), oc.keys())
        for k in rc:
            assert rc[k].shape == oc[k].shape and rc[k].dtype == oc[k].dtype, (t, k, rc[k].shape, oc[k].shape)
            assert torch.allclose(rc[k], oc[k], rtol=0, atol=1e-6), (t, k, float((rc[k] - oc[k]).abs().max()))
    return dev


TREE_SEED, TREE_ORDER = 6, [0, 4, 2, 5, 1, 3]
WILD = {"seed": 20, "async": True, "async_mode": "sim", "async_overhead": 100, "loc_err": True, "xyz_std": 0.2, "ryp_std": 0.2,
        "data_size": 0, "transmission_speed": 27, "backbone_delay": 0}


def tree_hypes(hypes, tree):
    """training on a directory with one timestamp of communication delay, localisation noise and the delayed ego pose"""
    h = copy.deepcopy(hypes)
    h.update(root_dir=tree, validate_dir=tree, task="det", wild_setting=dict(WILD))
    h["fusion"]["args"]["cur_ego_pose_flag"] = False
    return h


def run_tree(ds, order=TREE_ORDER, seed=3):
    import random
    random.seed(seed)
    np.random.seed(seed)
    items = [ds[i] for i in order]
    return items, ds.collate_batch_train(items)


def tree_case(IFD, DS, hypes):
    """the reference's FULL path — its own `__init__` directory scan, `retrieve_base_data`, `reform_param`, ego re-draw, time
    delay, localisation noise — on the seeded synthetic tree of tests/scenes_common.py, recorded for the machines that have
    no reference tree (open3d's reader replaced by this repo's `read_pcd`: the .pcd decoding itself stays unpinned)"""
    import shutil
    import tempfile
    from unittest.mock import MagicMock

    import scenes_common as SC
    from opencood.utils import pcd_utils
    S = a2x_import.pkg("airv2x_scenes")
    tmp = tempfile.mkdtemp(prefix="a2x_tree_")
    try:
        tree = SC.write_tree(os.path.join(tmp, "tree"), seed=TREE_SEED, late_agent=False)
        h = tree_hypes(hypes, tree)

        def fake_read(path):
            c = S.read_pcd(path)
            return MagicMock(points=c[:, :3].astype(np.float64), colors=np.stack([c[:, 3]] * 3, axis=1).astype(np.float64))
        pcd_utils.o3d.io.read_point_cloud = fake_read
        ref = IFD.IntermediateFusionDatasetAirv2x(h, False, True)
        ref.pre_processor = _OraclePreprocessor(h["preprocess"], True)
        ref_items, ref_batch = run_tree(ref)
        mine = DS.IntermediateFusionDatasetAirv2x(h, False, True, source=S.AirV2XScenes(h, True, load_cameras=True, load_seg=True))
        my_items, my_batch = run_tree(mine)
        dev = compare(ref_batch, my_batch, h, True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    r = ref_batch["ego"]
    print("tree: %d samples, egos %s, record_len %s, box deviation %.1e: reference == ours"
          % (len(ref_items), [a["ego"]["ego_id"] for a in ref_items], r["record_len"].tolist(), dev))
    out = {"tree/ego_ids": np.array([a["ego"]["ego_id"] for a in ref_items], dtype=np.int64),
           "tree/timestamp_keys": np.array([a["ego"]["timestamp_key"] for a in ref_items], dtype=np.int64),
           "tree/object_ids": np.array([i for ids in r["object_ids"] for i in ids], dtype=np.int64)}
    for k in ("record_len", "pairwise_t_matrix_collab", "prior_encoding", "spatial_correction_matrix", "object_bbx_center",
              "object_bbx_mask"):
        out["tree/" + k] = r[k].numpy()
    for t in ("vehicle", "rsu", "drone"):
        v = r[t]["batch_merged_lidar_features_torch"]
        out["tree/%s/record_len" % t] = r[t]["record_len"].numpy()
        out["tree/%s/voxel_coords" % t] = v["voxel_coords"].numpy().astype(np.int32)
        out["tree/%s/voxel_sum" % t] = v["voxel_features"].numpy().astype(np.float64).sum(axis=(1, 2))
    return out


def main(dst_dir=None):
    """dst_dir: where to write dataset.npz / dataset_config.json (default: tests/golden)"""
    dst_dir = dst_dir or os.path.join(ROOT, "tests", "golden")
    IFD = reference_env()
    DS = a2x_import.pkg("intermediate_fusion_dataset")
    import dataset_common as DC
    hypes = ref_import.load_hypes(YAML)
    out = {}
    for name, (train, specs) in CASES.items():
        scenes = [DC.synth_scene(DS, **kw) for kw in specs]
        ref_ds = reference_dataset(IFD, hypes, train)
        _, ref_batch = run_reference(ref_ds, scenes, seed=5)
        _, _, ours_batch = run_ours(DS, hypes, train, scenes, seed=5)
        dev = compare(ref_batch, ours_batch, hypes, train)
        r = ref_batch["ego"]
        print("%s: %d scenes, record_len %s, %d boxes, box deviation %.1e: reference == ours"
              % (name, len(scenes), r["record_len"].tolist(), int(r["object_bbx_mask"].sum()), dev))
        for k in ("record_len", "pairwise_t_matrix_collab", "img_pairwise_t_matrix_collab", "prior_encoding",
                  "spatial_correction_matrix", "object_bbx_center", "object_bbx_mask"):
            out["%s/%s" % (name, k)] = r[k].numpy()
        out[name + "/object_ids"] = np.array([i for ids in r["object_ids"] for i in ids], dtype=np.int64)
        out[name + "/class_ids"] = np.array([i for ids in r["class_ids"] for i in ids], dtype=np.int64)
        pos = r["label_dict"]["pos_equal_one"].numpy()
        out[name + "/pos_idx"] = np.flatnonzero(pos.reshape(-1))
        for t in ("vehicle", "rsu", "drone"):
            out["%s/%s/record_len" % (name, t)] = r[t]["record_len"].numpy()
            out["%s/%s/batch_idxs" % (name, t)] = np.array(r[t]["batch_idxs"], dtype=np.int64)
            v = r[t]["batch_merged_lidar_features_torch"]
            if v is not None:       # a digest of the voxel tensors: coordinates + point counts + per-voxel feature sums
                out["%s/%s/voxel_coords" % (name, t)] = v["voxel_coords"].numpy().astype(np.int32)
                out["%s/%s/voxel_num_points" % (name, t)] = v["voxel_num_points"].numpy().astype(np.int32)
                out["%s/%s/voxel_sum" % (name, t)] = v["voxel_features"].numpy().astype(np.float64).sum(axis=(1, 2))
            for k, c in r[t]["batch_merged_cam_inputs"].items():
                if k != "imgs":
                    out["%s/%s/cam_%s" % (name, t, k)] = c.numpy()
                else:
                    out["%s/%s/cam_imgs_mean" % (name, t)] = c.numpy().astype(np.float64).mean(axis=(2, 3, 4))
    out.update(tree_case(IFD, DS, hypes))
    import json
    keys = ("fusion", "preprocess", "postprocess", "train_params", "collaborators", "active_sensors", "ego_type")
    with open(os.path.join(dst_dir, "dataset_config.json"), "w") as f:   # the yaml keys the dataset reads
        json.dump({k: hypes[k] for k in keys}, f, default=lambda o: o.tolist())
    dst = os.path.join(dst_dir, "dataset.npz")
    np.savez_compressed(dst, **out)
    print("wrote %s (%.0f kB)" % (dst, os.path.getsize(dst) / 1e3))


if __name__ == "__main__":
    main()
