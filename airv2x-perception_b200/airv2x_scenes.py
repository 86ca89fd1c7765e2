"""Disk side of the dataset: the AirV2X directory scan of `BaseDataset.__init__` (basedataset.py:73-207,
`parse_seq` utils/airv2x_utils.py:121-263) and `retrieve_base_data` / `reform_param` / `calc_dist_to_ego` /
`time_delay_calculation` (basedataset.py:217-303, :305-532, :551-596, :727-758) as a scene SOURCE for
`intermediate_fusion_dataset.IntermediateFusionDatasetAirv2x(params, ..., source=AirV2XScenes(params, train))`.

    <root_dir>/<scenario>/timestamp_%06d/objects.pkl
    <root_dir>/<scenario>/timestamp_%06d/agent_%06d/{metadata.pkl, lidar.pcd, *_camera.png, *_depth.png, map_*.png}

What differs from the reference, on purpose:
  * the metadata / object pickles are parsed once and kept (the reference re-reads the ego's pickles for every agent of every
    sample: 4 + 4 loads per agent and `__getitem__`);
  * camera / depth PNGs are decoded only when `"cam"` is an active sensor and the segmentation maps only for the
    segmentation task — a lidar detection sample then touches one .pcd and one pickle per agent;
  * an agent that does not exist at the sample's timestamp is skipped (the reference drops into pdb, :576-586); entries of
    the tree that are not `<word>_<number>` folders, and agent folders without a metadata.pkl, are ignored by the scan;
  * .pcd files are read by `read_pcd` below (open3d, which the reference calls, is not a dependency here): x, y, z and the
    first colour channel as intensity, as `pcd_to_np` returns them (utils/pcd_utils.py:43-82). Parity of this reader with
    open3d is UNPINNED (open3d absent in the build container); everything else is pinned live against the reference class
    on a synthetic tree (tests/test_scenes_cpu.py).

The agent order inside a scenario is state, like the reference's `scenario_database`: agents that cannot be the ego are
moved behind the first ego-type agent at scan time, and in training every retrieval re-draws the ego of EVERY scenario from
Python's `random` and moves it to the front (`shuffle_ego`, :534-549) — same draws, same order. Host logic only.
"""
import math
import os
import pickle
import random
import re
from collections import OrderedDict

import numpy as np

from .intermediate_fusion_dataset import abs_world_pose, relative_pose_matrix

_CAMERAS = {"vehicle": ["front", "front_left", "front_right", "rear", "rear_left", "rear_right"],
            "rsu": ["back", "front", "left", "right"], "drone": ["bev"]}            # utils/airv2x_utils.py:37-118
_STATIC_MAPS = ["map_static_background.png", "map_static_lane.png", "map_static_road.png"]
_DYNAMIC_MAPS = ["map_dynamic_bev_layer_%d.png" % i for i in range(7)]
_INDEXED = re.compile(r"^[A-Za-z]+_\d+$")          # timestamp_000010, agent_001514
_PCD_TYPES = {("F", 4): "f4", ("F", 8): "f8", ("U", 1): "u1", ("U", 2): "u2", ("U", 4): "u4", ("I", 1): "i1",
              ("I", 2): "i2", ("I", 4): "i4"}


def lzf_decompress(src, out_len):
    """LZF (liblzf, the codec of PCD `DATA binary_compressed`): control byte < 32 = literal run of ctrl + 1 bytes; otherwise
    a back reference of length (ctrl >> 5) + 2 (7 = extended by the next byte) at distance ((ctrl & 31) << 8 | next) + 1."""
    src = memoryview(src)
    out = bytearray(out_len)
    i, o, n = 0, 0, len(src)
    while i < n:
        ctrl = src[i]
        i += 1
        if ctrl < 32:
            run = ctrl + 1
            if i + run > n or o + run > out_len:
                raise ValueError("corrupt LZF stream")
            out[o:o + run] = src[i:i + run]
            i += run
            o += run
            continue
        length = ctrl >> 5
        if i + (2 if length == 7 else 1) > n:
            raise ValueError("corrupt LZF stream")
        if length == 7:
            length += src[i]
            i += 1
        length += 2
        ref = o - (((ctrl & 31) << 8) | src[i]) - 1
        i += 1
        if ref < 0 or o + length > out_len:
            raise ValueError("corrupt LZF stream")
        if ref + length <= o:
            out[o:o + length] = out[ref:ref + length]
        else:                                   # overlapping copy = a repeating pattern of period o - ref
            period = o - ref
            chunk = bytes(out[ref:o])
            out[o:o + length] = (chunk * (length // period + 1))[:length]
        o += length
    if o != out_len:
        raise ValueError("LZF stream ends at %d of %d bytes" % (o, out_len))
    return bytes(out)


def read_pcd(path):
    """PCD v0.7 (ascii / binary / binary_compressed) -> [n, 4] float32 (x, y, z, intensity). Intensity = red channel / 255 of the packed `rgb`
    field (how open3d exposes `colors[:, 0]` for the clouds the simulator saved), or an `intensity` field if present."""
    with open(path, "rb") as f:
        fields, sizes, types, counts, n, mode, wh = [], [], [], [], None, None, [None, 1]
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: PCD header without DATA line" % path)
            tok = line.decode("ascii", "replace").strip().split()
            if not tok or tok[0].startswith("#"):
                continue
            key = tok[0].upper()
            if key == "FIELDS":
                fields = tok[1:]
            elif key == "SIZE":
                sizes = [int(v) for v in tok[1:]]
            elif key == "TYPE":
                types = tok[1:]
            elif key == "COUNT":
                counts = [int(v) for v in tok[1:]]
            elif key == "POINTS":
                n = int(tok[1])
            elif key == "WIDTH":
                wh[0] = int(tok[1])
            elif key == "HEIGHT":
                wh[1] = int(tok[1])
            elif key == "DATA":
                mode = tok[1].lower()
                break
        if n is None:                            # POINTS is optional in old files: WIDTH x HEIGHT
            if wh[0] is None:
                raise ValueError("%s: PCD header without POINTS / WIDTH" % path)
            n = wh[0] * wh[1]
        if not fields or len(sizes) != len(fields) or len(types) != len(fields) or any(a not in fields for a in "xyz"):
            raise ValueError("%s: PCD header needs FIELDS x y z with SIZE / TYPE entries" % path)
        counts = counts or [1] * len(fields)
        if any(c != 1 for c in counts):
            raise NotImplementedError("%s: multi-count PCD fields" % path)
        dt = np.dtype([(fields[i], _PCD_TYPES[(types[i].upper(), sizes[i])]) for i in range(len(fields))])
        if mode == "binary":
            rec = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
        elif mode == "ascii":
            raw = np.loadtxt(f, dtype=np.float64, ndmin=2)[:n]
            rec = np.zeros(raw.shape[0], dtype=dt)
            for i, name in enumerate(fields):        # a float `rgb` keeps its bit pattern: 10 significant digits round-trip fp32
                rec[name] = raw[:, i].astype(dt[name])
        elif mode == "binary_compressed":     # two uint32 sizes, LZF payload, fields stored one after the other (SoA)
            csize, usize = np.frombuffer(f.read(8), dtype="<u4")
            flat = lzf_decompress(f.read(int(csize)), int(usize))
            rec = np.zeros(n, dtype=dt)
            pos = 0
            for name in fields:
                rec[name] = np.frombuffer(flat, dtype=dt[name], count=n, offset=pos)
                pos += n * dt[name].itemsize
        else:
            raise NotImplementedError("%s: PCD DATA %s" % (path, mode))
    out = np.zeros((rec.shape[0], 4), dtype=np.float32)
    for i, name in enumerate("xyz"):
        out[:, i] = rec[name]
    if "rgb" in fields:
        packed = rec["rgb"].astype(np.float32).view(np.uint32) if rec["rgb"].dtype.kind == "f" else rec["rgb"].astype(np.uint32)
        out[:, 3] = ((packed >> 16) & 255).astype(np.float32) / 255.0
    elif "intensity" in fields:
        out[:, 3] = rec["intensity"]
    return out


def segmentation_label(map_files, kind):
    """binary PNG layers -> [W, H] uint8 class map, later layers on top, transposed and flipped into the lidar frame
    (`_wrap_segmentation_map`, basedataset.py:885-912)"""
    from PIL import Image
    assert len(map_files) == (7 if kind == "dynamic" else 3), "%s segmentation needs %d layers" % (kind, 7 if kind == "dynamic" else 3)
    layers = np.array([(np.array(Image.open(f).convert("L")) > 10).astype(np.uint8) for f in map_files])
    label = np.zeros(layers.shape[1:], dtype=np.uint8)
    for c in range(layers.shape[0]):
        label[layers[c] == 1] = c
    return label.T[:, ::-1]


class AirV2XScenes:
    """`len(src)` samples; `src[idx]` -> `(base_data_dict, scenario_index, timestamp_key)` = `retrieve_base_data(idx)`."""

    def __init__(self, params, train=True, load_cameras=None, load_seg=None, read_cloud=read_pcd):
        self.params, self.train, self.read_cloud = params, train, read_cloud
        self.load_cameras = ("cam" in params.get("active_sensors", [])) if load_cameras is None else load_cameras
        self.load_seg = (params.get("task", "det") != "det") if load_seg is None else load_seg
        ws = params.get("wild_setting")
        self.async_flag = bool(ws["async"]) if ws else False
        self.async_mode = ws.get("async_mode", "sim") if ws else "sim"
        self.async_overhead = ws["async_overhead"] if ws else 0
        self.loc_err_flag = bool(ws["loc_err"]) if ws else False
        self.xyz_noise_std, self.ryp_noise_std = (ws["xyz_std"], ws["ryp_std"]) if ws else (0, 0)
        self.seed = ws["seed"] if ws else None
        self.data_size = ws.get("data_size", 0) if ws else 0
        self.transmission_speed = ws.get("transmission_speed", 27) if ws else 27
        self.backbone_delay = ws.get("backbone_delay", 0) if ws else 0
        self.cur_ego_pose_flag = params["fusion"]["args"].get("cur_ego_pose_flag", True)
        self.correct_lidar_coordinate_system = params.get("correct_lidar_coordinate_system", False)
        self.ego_type = params.get("ego_type", "vehicle")
        root = params["root_dir"] if train else params["validate_dir"]
        self.scenarios, self.order, self.len_record = [], [], []
        self._pickles = {}
        for name in sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d))):
            agents = self._scan(os.path.join(root, name))
            order = list(agents.keys())
            if not any(self._type(agents[a]) == self.ego_type for a in order):
                raise ValueError("scenario %s holds no %s that could be the ego" % (name, self.ego_type))
            while self._type(agents[order[0]]) != self.ego_type:         # non-ego types behind the first ego-type agent
                order.append(order.pop(0))
            self.scenarios.append(agents)
            self.order.append(order)
            self.len_record.append((self.len_record[-1] if self.len_record else 0) + len(agents[order[0]]))
        self.ego = [None] * len(self.scenarios)

    # -- scan ----------------------------------------------------------------------------------------------
    @staticmethod
    def _type(agent):
        return next(iter(agent.values()))["agent_type"]

    def _pickle(self, path):
        v = self._pickles.get(path)
        if v is None:
            with open(path, "rb") as f:
                v = self._pickles[path] = pickle.load(f)
        return v

    def _scan(self, folder):
        """agent id -> timestamp id -> file record, agents in first-seen order (`parse_seq` + `convert2opv2v`)"""
        agents = OrderedDict()
        for ts_path in sorted(os.path.join(folder, t) for t in os.listdir(folder)):
            if not os.path.isdir(ts_path) or not _INDEXED.match(os.path.basename(ts_path)):
                continue                          # stray files / folders (the reference's int(name.split("_")[1]) would raise)
            ts = int(os.path.basename(ts_path).split("_")[1])
            for a_path in sorted(os.path.join(ts_path, a) for a in os.listdir(ts_path)):
                meta_path = os.path.join(a_path, "metadata.pkl")
                if os.path.isfile(a_path) or not _INDEXED.match(os.path.basename(a_path)) or not os.path.isfile(meta_path):
                    continue
                kind = self._pickle(meta_path)["agent_type"]
                if kind not in _CAMERAS:
                    raise ValueError("Unknown agent type: %s" % kind)
                have = lambda names: [os.path.join(a_path, n) for n in names if os.path.isfile(os.path.join(a_path, n))]  # noqa: E731
                rec = {"agent_type": kind, "metadata_path": meta_path, "objects": os.path.join(ts_path, "objects.pkl"),
                       "cameras": have([c + "_camera.png" for c in _CAMERAS[kind]]),
                       "depth": have([c + "_depth.png" for c in _CAMERAS[kind]]),
                       "lidars": have(["lidar.pcd", "semantic_lidar.pcd", "semantic_lidar_semantic.npz"]),
                       "map": have(["vector_map.json"] + _STATIC_MAPS + _DYNAMIC_MAPS)}
                agents.setdefault(int(os.path.basename(a_path).split("_")[1]), OrderedDict())[ts] = rec
        return agents

    def __len__(self):
        return self.len_record[-1] if self.len_record else 0

    # -- one sample -------------------------------------------------------------------------------------------
    def _shuffle_ego(self):
        for s, agents in enumerate(self.scenarios):
            order = self.order[s]
            if self.train:
                ego = random.choice([a for a in order if self._type(agents[a]) == self.ego_type])
                order.remove(ego)
                order.insert(0, ego)
            self.ego[s] = order[0]

    def _time_delay(self, is_ego):
        if is_ego:
            return 0
        if self.async_mode == "real":
            delay = int(np.random.uniform(0, self.async_overhead) + self.data_size / self.transmission_speed * 1000
                        + self.backbone_delay)
        else:
            delay = np.abs(self.async_overhead)
        return (delay // 100) if self.async_flag else 0

    def _noisy(self, pose):
        """`add_loc_noise` (:699-725): re-seeds numpy's global generator on every call, yaw noise only"""
        np.random.seed(self.seed)
        xyz = np.random.normal(0, self.xyz_noise_std, 3)
        ryp = np.random.normal(0, self.ryp_noise_std, 3)
        return [pose[0] + xyz[0], pose[1] + xyz[1], pose[2] + xyz[2], pose[3], pose[4] + ryp[1], pose[5]]

    def _params(self, agent, ego_agent, is_ego, ts_cur, ts_delay):
        """`reform_param`: the delayed metadata plus poses, camera calibration and the CURRENT objects (classes 1..6)"""
        cur, delay = self._pickle(agent[ts_cur]["metadata_path"]), self._pickle(agent[ts_delay]["metadata_path"])
        ego_cur, ego_delay = (self._pickle(ego_agent[ts_cur]["metadata_path"]),
                              self._pickle(ego_agent[ts_delay]["metadata_path"]))
        lidar = lambda m: abs_world_pose(m["lidar"]["lidar_pose"], m["odometry"]["ego_pos"])  # noqa: E731
        cur_ego, delay_ego = lidar(ego_cur), lidar(ego_delay)
        delay_cav, cur_cav = lidar(delay), lidar(cur)
        if not is_ego and self.loc_err_flag:
            delay_cav, cur_cav = self._noisy(delay_cav), self._noisy(cur_cav)
        out = dict(delay)
        out.update(cur_ego_lidar_pose=cur_ego, delay_ego_lidar_pose=delay_ego, cur_cav_lidar_pose=cur_cav,
                   delay_cav_lidar_pose=delay_cav)
        if self.cur_ego_pose_flag:
            out["transformation_matrix"], out["spatial_correction_matrix"] = relative_pose_matrix(delay_cav, cur_ego), np.eye(4)
        else:
            out["transformation_matrix"] = relative_pose_matrix(delay_cav, delay_ego)
            out["spatial_correction_matrix"] = relative_pose_matrix(delay_ego, cur_ego)
        out["gt_transformation_matrix"] = relative_pose_matrix(cur_cav, cur_ego)
        for tag, meta in (("cur", cur), ("delay", delay)):
            cams = [c + "_camera" for c in _CAMERAS[meta["agent_type"]]]
            out[tag + "_intrinsic"] = np.array([np.array(meta[c]["intrinsic"], dtype=np.float32) for c in cams])
            out[tag + "_extrinsic"] = np.array([np.array(meta[c]["extrinsic"], dtype=np.float32) for c in cams])
        out["delay_lidar_ego_abs_pos"], out["cur_lidar_ego_abs_pos"] = delay_ego, cur_ego
        objs = self._pickle(agent[ts_cur]["objects"])
        out["objects"] = {k: o for k, o in objs.items() if o["class"] in (1, 2, 3, 4, 5, 6)}
        return out

    def __getitem__(self, idx):
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError("sample %d of %d" % (idx, len(self)))
        s = next(i for i, end in enumerate(self.len_record) if idx < end)
        agents = self.scenarios[s]
        t_index = idx if s == 0 else idx - self.len_record[s - 1]
        ts_cur = list(agents[self.order[s][0]].keys())[t_index]          # read off the first agent BEFORE the ego re-draw
        self._shuffle_ego()
        ego_id = self.ego[s]
        ego_agent = agents[ego_id]
        ego_pos = self._pickle(ego_agent[ts_cur]["metadata_path"])["odometry"]["ego_pos"]
        data = OrderedDict()
        for a in self.order[s]:
            agent = agents[a]
            if ts_cur not in agent:
                continue
            pos = self._pickle(agent[ts_cur]["metadata_path"])["odometry"]["ego_pos"]
            delay = self._time_delay(a == ego_id)
            if t_index - delay <= 0:
                delay = t_index
            ts_delay = list(ego_agent.keys())[max(0, t_index - delay)]
            if ts_delay not in agent:
                ts_delay = ts_cur
            files = agent[ts_delay]
            rec = OrderedDict(ego=(a == ego_id), agent_type=agent[ts_cur]["agent_type"],
                              distance_to_ego=math.sqrt(sum((pos[i] - ego_pos[i]) ** 2 for i in range(3))),
                              time_delay=delay)
            rec["params"] = self._params(agent, ego_agent, a == ego_id, ts_cur, ts_delay)
            rec["cameras"], rec["depth"] = [], []
            if self.load_cameras:
                from PIL import Image
                rec["cameras"] = [Image.open(f).copy() for f in files["cameras"]]
                rec["depth"] = [Image.open(f).copy() for f in files["depth"]]
            cloud = self.read_cloud(files["lidars"][0])
            if self.correct_lidar_coordinate_system:
                cloud[:, 1] = -cloud[:, 1]
            rec["lidar_np"] = cloud
            if self.load_seg:
                rec["dynamic_seg_label"] = segmentation_label(files["map"][-7:], "dynamic")
                rec["static_seg_label"] = segmentation_label(files["map"][-10:-7], "static")
            rec["metadata_path"] = files["metadata_path"]
            data[a] = rec
        return data, s, ts_cur
