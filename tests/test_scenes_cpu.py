"""CPU (-m "not gpu"): the disk side of the dataset (`airv2x-perception_b200/airv2x_scenes.py`).

* everywhere: the PCD reader round-trips the writer of the test tree (ascii and binary), the scan orders agents / counts
  samples as documented, an agent missing at a timestamp is skipped, the dataset built from a directory collates;
* where /root/reference exists: the REAL `IntermediateFusionDatasetAirv2x(params, visualize, train)` — its own `__init__`
  directory scan, `retrieve_base_data`, `reform_param`, ego shuffling, time delays — runs on the same synthetic tree and
  both datasets must emit the same batches sample for sample. The reference's open3d reader (absent here) is replaced by
  this repo's `read_pcd` on the reference side, so the .pcd decoding itself stays unpinned (see the module docstring)."""
import copy
import os
import random
import sys
from unittest.mock import MagicMock

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import a2x_import  # noqa: E402
import make_golden_dataset as MGD  # noqa: E402
import scenes_common as SC  # noqa: E402
from oracle import ref_import  # noqa: E402

needs_reference = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    return SC.write_tree(str(tmp_path_factory.mktemp("airv2x")), seed=4)


@pytest.fixture(scope="module")
def hypes(tree):
    import json
    h = json.load(open(os.path.join(ROOT, "tests", "golden", "dataset_config.json")))
    h.update(root_dir=tree, validate_dir=tree, task="det")
    return h


def test_pcd_reader_round_trips_binary_and_ascii(tmp_path):
    S = a2x_import.pkg("airv2x_scenes")
    g = np.random.default_rng(1)
    cloud = np.concatenate([g.normal(0, 30, (257, 3)), g.integers(0, 256, (257, 1)) / 255.0], axis=1).astype(np.float32)
    SC.write_pcd(str(tmp_path / "b.pcd"), cloud)
    got = S.read_pcd(str(tmp_path / "b.pcd"))
    assert got.dtype == np.float32 and np.array_equal(got, cloud)
    with open(tmp_path / "a.pcd", "w") as f:
        f.write("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 3\nHEIGHT 1\n"
                "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS 3\nDATA ascii\n1 2 3 0.5\n-4.5 5 6 0.25\n7 8 -9 1\n")
    assert np.array_equal(S.read_pcd(str(tmp_path / "a.pcd")),
                          np.array([[1, 2, 3, 0.5], [-4.5, 5, 6, 0.25], [7, 8, -9, 1]], dtype=np.float32))
    # ascii with the packed float rgb (denormals for r < 128) and with an unsigned rgb
    r = np.array([0, 1, 127, 128, 255], dtype=np.uint32)
    for kind, col in (("F", ["%.10g" % v for v in (r << 16).view(np.float32)]), ("U", [str(int(v)) for v in (r << 16) | 0x1234])):
        with open(tmp_path / "c.pcd", "w") as f:
            f.write("VERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F %s\nCOUNT 1 1 1 1\nWIDTH 5\nHEIGHT 1\nPOINTS 5\nDATA ascii\n" % kind)
            for i in range(5):
                f.write("%d 0 0 %s\n" % (i, col[i]))
        assert np.array_equal(S.read_pcd(str(tmp_path / "c.pcd"))[:, 3], r.astype(np.float32) / 255.0), kind
    # binary_compressed (LZF): a cloud with repeated values (back references, overlapping runs) and the decoder alone
    rep = np.repeat(cloud[:40], 5, axis=0)
    rep[:, 2] = 0.25
    SC.write_pcd_compressed(str(tmp_path / "z.pcd"), rep)
    assert np.array_equal(S.read_pcd(str(tmp_path / "z.pcd")), rep)
    for blob in (b"", b"a", b"abcabcabcabcabcabcabcabcabc" * 20, bytes(300), bytes(g.integers(0, 4, 5000, dtype=np.uint8)),
                 bytes(g.integers(0, 256, 3000, dtype=np.uint8))):
        comp = SC.lzf_compress(blob)
        assert S.lzf_decompress(comp, len(blob)) == blob
        if len(blob) >= 300 and len(set(blob)) <= 4:
            assert len(comp) < len(blob) // 2                    # the encoder really emits back references
    with pytest.raises(ValueError):
        S.lzf_decompress(b"\x20\x05", 10)                       # reference before the start of the output
    with open(tmp_path / "w.pcd", "w") as f:                      # no POINTS line: WIDTH x HEIGHT; extra fields are carried along
        f.write("VERSION 0.7\nFIELDS x y z normal_x intensity\nSIZE 4 4 4 4 4\nTYPE F F F F F\nWIDTH 2\nHEIGHT 1\nDATA ascii\n"
                "1 2 3 9 0.5\n4 5 6 9 0.75\n")
    assert np.array_equal(S.read_pcd(str(tmp_path / "w.pcd")), np.array([[1, 2, 3, 0.5], [4, 5, 6, 0.75]], dtype=np.float32))
    for bad in ("VERSION 0.7\nFIELDS x y\nSIZE 4 4\nTYPE F F\nPOINTS 1\nDATA ascii\n1 2\n",
                "VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nDATA ascii\n1 2 3\n"):
        with open(tmp_path / "bad.pcd", "w") as f:
            f.write(bad)
        with pytest.raises(ValueError):
            S.read_pcd(str(tmp_path / "bad.pcd"))
    with open(tmp_path / "q.pcd", "w") as f:
        f.write("VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA quantum\n")
    with pytest.raises(NotImplementedError):
        S.read_pcd(str(tmp_path / "q.pcd"))


def test_scan_orders_agents_and_counts_samples(hypes):
    S = a2x_import.pkg("airv2x_scenes")
    first = sorted(os.listdir(hypes["root_dir"]))[0]                    # stray entries in the tree are ignored by the scan
    os.makedirs(os.path.join(hypes["root_dir"], first, "logs"), exist_ok=True)
    os.makedirs(os.path.join(hypes["root_dir"], first, "timestamp_000000", "agent_000099"), exist_ok=True)   # no metadata
    open(os.path.join(hypes["root_dir"], first, "notes.txt"), "w").close()
    src = S.AirV2XScenes(hypes, train=False)
    assert len(src) == 6 and src.len_record == [3, 6]
    # ids sorted by path, the leading RSU moved behind the first vehicle; vehicle 35 appears at the 2nd timestamp in scenario 0
    assert src.order[0] == [11, 12, 20, 21, 35, 3] and src.order[1] == [11, 12, 20, 21, 35, 3]
    base, s, ts = src[0]
    assert (s, ts) == (0, 0) and list(base.keys()) == [11, 12, 20, 21, 3] and base[11]["ego"] and not base[12]["ego"]
    assert base[11]["distance_to_ego"] == 0.0 and base[3]["agent_type"] == "rsu"
    assert base[11]["cameras"] == [] and "dynamic_seg_label" not in base[11]          # lidar detection: PNGs not decoded
    assert base[11]["lidar_np"].shape == (600, 4) and set(o["class"] for o in base[11]["params"]["objects"].values()) <= {1, 2, 3, 4, 5, 6}
    base, s, ts = src[4]
    assert (s, ts) == (1, 5) and list(base.keys()) == [11, 12, 20, 21, 35, 3]
    full = S.AirV2XScenes(hypes, train=False, load_cameras=True, load_seg=True)[1][0]
    assert len(full[11]["cameras"]) == 6 and len(full[3]["depth"]) == 4 and len(full[12]["cameras"]) == 1
    assert full[11]["dynamic_seg_label"].shape == (24, 16) and full[11]["static_seg_label"].max() <= 2


def test_dataset_scans_the_directory_when_no_source_is_given(hypes):
    DS = a2x_import.pkg("intermediate_fusion_dataset")
    ds = DS.IntermediateFusionDatasetAirv2x(hypes, False, train=False)
    assert len(ds) == 6
    batch = ds.collate_batch_train([ds[0], ds[5]])["ego"]
    assert batch["record_len"].tolist() == [5, 6] and batch["raw_points"]["offsets"][-1] == 11 * 600
    assert batch["vehicle"]["record_len"].tolist() == [2, 3] and batch["drone"]["batch_idxs"] == [0, 1]


def test_directory_pipeline_equals_the_recorded_reference_run(hypes, tmp_path):
    """runs everywhere: this repo's scan + retrieval + assembly on the seeded tree against what the REAL reference (its own
    `__init__` scan, `retrieve_base_data`, `reform_param`; training mode, one timestamp of delay, localisation noise, delayed
    ego pose) produced for the same tree when scripts/make_golden_dataset.py recorded it"""
    S = a2x_import.pkg("airv2x_scenes")
    DS = a2x_import.pkg("intermediate_fusion_dataset")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "dataset.npz"))
    tree = SC.write_tree(str(tmp_path / "tree"), seed=MGD.TREE_SEED, late_agent=False)
    h = MGD.tree_hypes(hypes, tree)
    ds = DS.IntermediateFusionDatasetAirv2x(h, False, True, source=S.AirV2XScenes(h, True, load_cameras=True, load_seg=True))
    items, batch = MGD.run_tree(ds)
    o = batch["ego"]
    assert [a["ego"]["ego_id"] for a in items] == gold["tree/ego_ids"].tolist()
    assert [a["ego"]["timestamp_key"] for a in items] == gold["tree/timestamp_keys"].tolist()
    assert len(set(gold["tree/ego_ids"].tolist())) > 1                      # the ego really was re-drawn
    assert [i for ids in o["object_ids"] for i in ids] == gold["tree/object_ids"].tolist()
    for k in ("record_len", "prior_encoding", "object_bbx_mask"):
        assert np.array_equal(o[k].numpy(), gold["tree/" + k]), k
    assert float(o["prior_encoding"][:, 1:, 1].max()) == 1.0                # the delayed agents carry their time delay
    for k, tol in (("pairwise_t_matrix_collab", 1e-6), ("spatial_correction_matrix", 1e-12), ("object_bbx_center", 1e-9)):
        assert np.abs(o[k].numpy() - gold["tree/" + k]).max() <= tol, k
    corr = gold["tree/spatial_correction_matrix"]
    assert np.allclose(corr[:, 0], np.eye(4)) and not np.allclose(corr[:, 1:6], np.eye(4))   # delayed ego pose: real corrections
    vox = MGD.voxelise_like_the_reference(o, h, True)
    for t in ("vehicle", "rsu", "drone"):
        assert np.array_equal(o[t]["record_len"].numpy(), gold["tree/%s/record_len" % t])
        assert np.array_equal(vox[t]["voxel_coords"], gold["tree/%s/voxel_coords" % t])
        assert np.allclose(vox[t]["voxel_features"].astype(np.float64).sum(axis=(1, 2)), gold["tree/%s/voxel_sum" % t],
                           rtol=1e-6, atol=1e-3)


@needs_reference
@pytest.mark.parametrize("train,delay", [(False, False), (True, False), (True, True)])
def test_live_against_the_reference_dataset_on_a_directory(hypes, train, delay, tmp_path, monkeypatch):
    import pdb

    def no_pdb(*a, **k):
        raise RuntimeError("the reference dropped into pdb")
    monkeypatch.setattr(pdb, "set_trace", no_pdb)
    monkeypatch.chdir(tmp_path)
    # every agent present at every timestamp: the reference stops in pdb otherwise (basedataset.py:576-586)
    tree = SC.write_tree(str(tmp_path / "tree"), seed=6, late_agent=False)
    hypes = dict(hypes, root_dir=tree, validate_dir=tree)
    IFD = MGD.reference_env()
    S = a2x_import.pkg("airv2x_scenes")
    DS = a2x_import.pkg("intermediate_fusion_dataset")
    full = ref_import.load_hypes(MGD.YAML)
    full.update(root_dir=hypes["root_dir"], validate_dir=hypes["validate_dir"])
    if delay:
        full["wild_setting"] = {"seed": 20, "async": True, "async_mode": "sim", "async_overhead": 100, "loc_err": True,
                                "xyz_std": 0.2, "ryp_std": 0.2, "data_size": 0, "transmission_speed": 27, "backbone_delay": 0}
        full["fusion"]["args"]["cur_ego_pose_flag"] = False
    from opencood.utils import pcd_utils

    def fake_read(path):      # open3d is absent: the reference's pcd_to_np gets .points / .colors from this repo's reader
        c = S.read_pcd(path)
        return MagicMock(points=c[:, :3].astype(np.float64), colors=np.stack([c[:, 3]] * 3, axis=1).astype(np.float64))
    monkeypatch.setattr(pcd_utils.o3d.io, "read_point_cloud", fake_read, raising=False)
    ref = IFD.IntermediateFusionDatasetAirv2x(full, False, train)          # the REAL __init__: directory scan and all
    ref.pre_processor = MGD._OraclePreprocessor(full["preprocess"], train)
    mine_cfg = copy.deepcopy({k: full[k] for k in hypes if k in full})
    mine_cfg.update(root_dir=full["root_dir"], validate_dir=full["validate_dir"], task="det")
    if delay:
        mine_cfg["wild_setting"] = full["wild_setting"]
    src = S.AirV2XScenes(mine_cfg, train, load_cameras=True, load_seg=True)
    mine = DS.IntermediateFusionDatasetAirv2x(mine_cfg, False, train, source=src)
    assert len(ref) == len(mine) == 6 and ref.len_record == src.len_record
    order = [0, 4, 2, 5, 1, 3]
    random.seed(3)
    np.random.seed(3)
    ref_items = [ref[i] for i in order]
    random.seed(3)
    np.random.seed(3)
    my_items = [mine[i] for i in order]
    for a, b in zip(ref_items, my_items):
        assert a["ego"]["ego_id"] == b["ego"]["ego_id"] and a["ego"]["timestamp_key"] == b["ego"]["timestamp_key"]
        assert a["ego"]["scenario_index"] == b["ego"]["scenario_index"]
        assert np.array_equal(a["ego"]["dynamic_seg_label"], b["ego"]["dynamic_seg_label"])
        assert np.array_equal(a["ego"]["static_seg_label"], b["ego"]["static_seg_label"])
    assert MGD.compare(ref.collate_batch_train(ref_items), mine.collate_batch_train(my_items), full, train) < 1e-9
    if train:   # the ego is re-drawn per sample: more than one vehicle was the ego over the six samples
        assert len({a["ego"]["ego_id"] for a in ref_items}) > 1


@needs_reference
def test_real_build_dataset_dispatches_to_the_b200_dataset(hypes, tmp_path, monkeypatch):
    """the UNMODIFIED `build_dataset(hypes, visualize, train)` returns this repo's class after `a2x_import.install_dataset()`
    and that object scans the directory the yaml names"""
    monkeypatch.chdir(tmp_path)
    MGD.reference_env()
    from opencood.data_utils.datasets import build_dataset
    full = ref_import.load_hypes(MGD.YAML)
    full.update(root_dir=hypes["root_dir"], validate_dir=hypes["validate_dir"])
    prev = a2x_import.install_dataset()
    try:
        ds = build_dataset(full, visualize=False, train=False)
        assert type(ds).__module__.startswith("airv2x-perception_b200.") and type(ds).__name__ == "IntermediateFusionDatasetAirv2x"
        assert len(ds) == 6 and ds.collate_batch_test([ds[1]])["ego"]["record_len"].tolist() == [6]
    finally:
        a2x_import.uninstall_dataset(prev)
    assert build_dataset.__module__ == "opencood.data_utils.datasets"
    import opencood.data_utils.datasets as D
    assert D.__all__["IntermediateFusionDatasetAirv2x"].__module__.startswith("opencood.")


def test_lzf_round_trip_property():
    """hypothesis: decode(encode(x)) == x for arbitrary byte strings (low-entropy ones exercise long / overlapping
    back references); truncated streams raise instead of returning short output"""
    from hypothesis import given, settings, strategies as st

    S = a2x_import.pkg("airv2x_scenes")

    @settings(max_examples=150, deadline=None)
    @given(st.one_of(st.binary(max_size=600), st.lists(st.sampled_from([0, 1, 255]), max_size=900).map(bytes)))
    def check(blob):
        comp = SC.lzf_compress(blob)
        assert S.lzf_decompress(comp, len(blob)) == blob
        if len(comp) > 1:
            with pytest.raises(ValueError):
                S.lzf_decompress(comp[:-1], len(blob))
    check()
