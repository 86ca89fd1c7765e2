"""Dataset side of the hot path: what `IntermediateFusionDatasetAirv2x.__getitem__` /
`get_item_single_car` / `collate_batch_train` (`opencood/data_utils/datasets/airv2x/intermediate_fusion_dataset.py:137-422,
:456-618, :620-888, :967-1020`) do between a loaded scene and `model(batch["ego"])`, arranged for the B200 path:

  * the per-agent clouds leave the worker RAW (shuffled, sensor frame, 16 B / point) together with each agent's
    `transformation_matrix`; the body-box filter, the projection to the ego frame, the range filter and the voxelisation
    (`get_item_single_car` :590-607, `SpVoxelPreprocessor`) run on the GPU (`a2x_voxelize_ex`), so a worker never builds the
    16 MB / agent padded voxel tensors and the batch carries `raw_points` instead of `batch_merged_lidar_features_torch`;
  * the anchor targets are assigned on the GPU (`labels.TargetAssigner`) from the padded ground-truth boxes this module
    emits (`object_bbx_center`, `object_bbx_mask`, `object_class_ids`), so no fp64 label maps cross PCIe;
  * everything else the models read — `record_len`, per-type `record_len` / `batch_idxs`, `pairwise_t_matrix_collab`,
    `img_pairwise_t_matrix_collab`, `prior_encoding`, `spatial_correction_matrix`, the camera geometry — is host
    bookkeeping and is reproduced value for value (tests/test_dataset_cpu.py drives the REAL reference class on the same
    scenes).

The boundary towards the disk is `retrieve_base_data(idx)` (`basedataset.py:217-303`): a `source` sequence / callable yields
the per-agent records it returns (`ego`, `agent_type`, `distance_to_ego`, `time_delay`, `params`, `lidar_np`, `cameras`);
without one the AirV2X directory tree under `params["root_dir"]` is scanned like the reference does (`airv2x_scenes.py`).
`agent_pose_params` restates the pose half of `reform_param` (`basedataset.py:305-532`) for in-memory sources.

Host logic only (numpy / torch CPU tensors): no kernel work happens here; the checker under `oracle/` is not used.
"""
import heapq
import math
import os
from collections import OrderedDict

import numpy as np
import torch

COM_RANGE = {"vehicle": 120, "rsu": 120, "drone": 180}      # data_utils/datasets/__init__.py:89-91
INFRA = {"vehicle": 0, "rsu": 1, "drone": 1}                # intermediate_fusion_dataset.py:160-176
ABBR = {"vehicle": "veh", "rsu": "rsu", "drone": "drone"}
MODEL_ORDER = ("vehicle", "rsu", "drone")                   # scene-major repack, airv2x_base_model.py:179-248
# corner signs of a box in its own frame (rows = corners 0..7), utils/box_utils.py:475-503
_CORNER_SIGNS = np.array([[1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1],
                          [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1]], dtype=np.float64)


# ----------------------------------------------------------------------------------------------------- poses
def pose_to_matrix(pose):
    """[x, y, z, roll, yaw, pitch] (degrees, CARLA convention) -> 4x4 float64 pose in the world frame
    (`x_to_world`, utils/transformation_utils.py:216-259)."""
    x, y, z, roll, yaw, pitch = [float(v) for v in pose[:6]]
    cy, sy = np.cos(np.radians(yaw)), np.sin(np.radians(yaw))
    cr, sr = np.cos(np.radians(roll)), np.sin(np.radians(roll))
    cp, sp = np.cos(np.radians(pitch)), np.sin(np.radians(pitch))
    m = np.identity(4)
    m[:3, 3] = (x, y, z)
    m[0, :3] = (cp * cy, cy * sp * sr - sy * cr, -cy * sp * cr - sy * sr)
    m[1, :3] = (sy * cp, sy * sp * sr + cy * cr, -sy * sp * cr + cy * sr)
    m[2, :3] = (sp, -cp * sr, cp * cr)
    return m


def relative_pose_matrix(src_pose, dst_pose):
    """T_dst<-src between two world poses (`x1_to_x2`, utils/transformation_utils.py:262-286)"""
    return np.dot(np.linalg.inv(pose_to_matrix(dst_pose)), pose_to_matrix(src_pose))


def abs_world_pose(rel_pose, center_pose):
    """a sensor pose given relative to its carrier (`get_abs_world_pose`, utils/transformation_utils.py:526-540)"""
    return [rel_pose[i] + center_pose[i] for i in range(6)]


def agent_pose_params(delay_meta, cur_meta, cur_ego_meta, delay_ego_meta, cur_ego_pose_flag=True):
    """The pose half of `reform_param` (basedataset.py:341-357, :450-492) for one agent: metadata dicts (`lidar.lidar_pose`,
    `odometry.ego_pos`) of the agent at the delayed / current timestamp and of the ego at both -> the entries of `params`
    the intermediate-fusion path reads."""
    lidar = lambda m: abs_world_pose(m["lidar"]["lidar_pose"], m["odometry"]["ego_pos"])  # noqa: E731
    cur_ego, delay_ego = lidar(cur_ego_meta), lidar(delay_ego_meta)
    delay_cav, cur_cav = lidar(delay_meta), lidar(cur_meta)
    if cur_ego_pose_flag:
        t, corr = relative_pose_matrix(delay_cav, cur_ego), np.eye(4)
    else:
        t, corr = relative_pose_matrix(delay_cav, delay_ego), relative_pose_matrix(delay_ego, cur_ego)
    return {"transformation_matrix": t, "spatial_correction_matrix": corr,
            "gt_transformation_matrix": relative_pose_matrix(cur_cav, cur_ego),
            "cur_ego_lidar_pose": cur_ego, "delay_ego_lidar_pose": delay_ego,
            "cur_cav_lidar_pose": cur_cav, "delay_cav_lidar_pose": delay_cav}


def distance_to_ego(agent_pos, ego_pos):
    """`calc_dist_to_ego` (basedataset.py:588-592): Euclidean distance of the odometry positions"""
    return math.sqrt(sum((agent_pos[i] - ego_pos[i]) ** 2 for i in range(3)))


def nearest_agents(distances, k):
    """positions (in input order) of the k entries of `[(cav_id, distance), ...]` with the smallest distance
    (`get_smallest_k_indices`, utils/airv2x_utils.py:495-522: heap selection, then the input order is kept)."""
    chosen = {cid for _, (cid, _) in heapq.nsmallest(k, enumerate(distances), key=lambda e: e[1][1])}
    idx = [i for i, (cid, _) in enumerate(distances) if cid in chosen][:k]
    return idx, [distances[i][0] for i in idx]


# ----------------------------------------------------------------------------------------------------- ground truth
def boxes_to_corners_f32(boxes7, order):
    """[n,7] centre boxes -> [n,8,3] corners in fp32, the arithmetic of `boxes_to_corners_3d`
    (utils/box_utils.py:195-257: `check_numpy_to_torch` casts to float, z-rotation as a row-vector matmul)."""
    b = torch.from_numpy(np.ascontiguousarray(boxes7)).float()
    if order == "hwl":
        b = b[:, [0, 1, 2, 5, 4, 3, 6]]
    half = torch.from_numpy(_CORNER_SIGNS).float() / 2
    c = b[:, None, 3:6].repeat(1, 8, 1) * half[None]
    cos, sin = torch.cos(b[:, 6]), torch.sin(b[:, 6])
    zero, one = torch.zeros_like(cos), torch.ones_like(cos)
    rot = torch.stack((cos, sin, zero, -sin, cos, zero, zero, zero, one), dim=1).view(-1, 3, 3)
    return (torch.matmul(c, rot) + b[:, None, 0:3]).numpy()


def corners_to_boxes(corners, order):
    """[n,8,3] corners -> [n,7] (x, y, z, then h,w,l or l,w,h, yaw): centre = mean of corners 0,3,5,6; every extent and
    the yaw are the mean over the four parallel edges (`corner_to_center`, utils/box_utils.py:28-133)."""
    c = np.asarray(corners, dtype=np.float64)
    xy = c[:, :, :2]
    edge = lambda a, b: np.sqrt(((xy[:, a] - xy[:, b]) ** 2).sum(axis=1, keepdims=True))  # noqa: E731
    ang = lambda a, b: np.arctan2(c[:, a, 1] - c[:, b, 1], c[:, a, 0] - c[:, b, 0])      # noqa: E731
    centre = np.mean(c[:, [0, 3, 5, 6], :], axis=1)
    h = abs(np.mean(c[:, 4:, 2] - c[:, :4, 2], axis=1, keepdims=True))
    l = (edge(0, 3) + edge(2, 1) + edge(4, 7) + edge(5, 6)) / 4
    w = (edge(0, 1) + edge(2, 3) + edge(4, 5) + edge(6, 7)) / 4
    yaw = (ang(1, 2) + ang(0, 3) + ang(5, 6) + ang(4, 7))[:, None] / 4
    dims = [l, w, h] if order == "lwh" else [h, w, l]
    return np.concatenate([centre] + dims + [yaw], axis=1).reshape(len(c), 7)


def project_world_objects(objects, ego_lidar_pose, lidar_range, order, cache=None):
    """World-frame objects `{id: {"location": [x,y,z,roll,yaw,pitch], "center", "extent", "class"}}` -> boxes in the ego
    lidar frame that lie inside `lidar_range` with all 8 corners: (boxes [n,7] f64, ids, classes), in dict order
    (`project_world_objects_airv2x`, utils/box_utils.py:576-647; the range test runs on fp32 corners like
    `mask_boxes_outside_range_numpy` :433-472).
    `cache` (a dict that lives for ONE ego pose / range, i.e. one scene): an object record that several agents list — the
    agents of a timestamp share one objects.pkl — is projected once; the per-object arithmetic does not depend on which
    other objects are in the call, so the values are those of the uncached evaluation."""
    ids = list(objects.keys())
    if not ids:
        return np.zeros((0, 7)), [], []
    cache = {} if cache is None else cache
    todo = [oid for oid in ids if id(objects[oid]) not in cache]
    if todo:
        world_to_ego = cache.get("world_to_ego")
        if world_to_ego is None:
            world_to_ego = cache["world_to_ego"] = np.linalg.inv(pose_to_matrix(ego_lidar_pose))
        corners = np.empty((len(todo), 8, 3))
        for i, oid in enumerate(todo):
            o = objects[oid]
            loc, ctr = o["location"], o["center"]
            pose = [loc[0] + ctr[0], loc[1] + ctr[1], loc[2] + ctr[2], loc[3], loc[4], loc[5]]
            local = np.r_[(_CORNER_SIGNS * np.asarray(o["extent"], dtype=np.float64)[None, :3]).T, [np.ones(8)]]   # [4,8]
            corners[i] = np.dot(np.dot(world_to_ego, pose_to_matrix(pose)), local).T[:, :3]
        boxes = corners_to_boxes(corners, order)
        c32 = boxes_to_corners_f32(boxes, order).astype(np.float64)
        lo, hi = np.asarray(lidar_range[0:3], dtype=np.float64), np.asarray(lidar_range[3:6], dtype=np.float64)
        keep = ((c32 >= lo) & (c32 <= hi)).all(axis=2).sum(axis=1) >= 8
        for i, oid in enumerate(todo):      # the record itself is kept alive next to its result: id() stays unique
            cache[id(objects[oid])] = (objects[oid], boxes[i], bool(keep[i]))
    sel = [oid for oid in ids if cache[id(objects[oid])][2]]
    out = np.stack([cache[id(objects[oid])][1] for oid in sel]) if sel else np.zeros((0, 7))
    return out, sel, [objects[oid]["class"] for oid in sel]


# ----------------------------------------------------------------------------------------------------- cameras
def sample_augmentation(conf, is_train):
    """resize / crop / flip / rotate of one camera (`sample_augmentation`, utils/camera_utils.py:31-57); the train branch
    draws from numpy's global generator in the reference's order, so a seeded run sees the same stream."""
    H, W = conf["H"], conf["W"]
    fH, fW = conf["final_dim"]
    if is_train:
        resize = np.random.uniform(*conf["resize_lim"])
        newW, newH = int(W * resize), int(H * resize)
        crop_h = int((1 - np.random.uniform(*conf["bot_pct_lim"])) * newH) - fH
        crop_w = int(np.random.uniform(0, max(0, newW - fW)))
        flip = bool(conf["rand_flip"] and np.random.choice([0, 1]))
        rotate = np.random.uniform(*conf["rot_lim"])
    else:
        resize = max(fH / H, fW / W)
        newW, newH = int(W * resize), int(H * resize)
        crop_h = int((1 - np.mean(conf["bot_pct_lim"])) * newH) - fH
        crop_w = int(max(0, newW - fW) / 2)
        flip, rotate = False, 0
    return resize, (newW, newH), (crop_w, crop_h, crop_w + fW, crop_h + fH), flip, rotate


def post_homography(resize, crop, flip, rotate):
    """the 3x3 `post_rot` / 3-vector `post_tran` an image augmentation induces on pixel coordinates
    (`img_transform`, utils/camera_utils.py:73-89, embedded in 3-D as intermediate_fusion_dataset.py:548-551)."""
    rot2 = torch.eye(2) * resize
    tran2 = torch.zeros(2) - torch.Tensor(crop[:2])
    if flip:
        A = torch.Tensor([[-1, 0], [0, 1]])
        rot2, tran2 = A.matmul(rot2), A.matmul(tran2) + torch.Tensor([crop[2] - crop[0], 0])
    h = rotate / 180 * np.pi
    A = torch.Tensor([[np.cos(h), np.sin(h)], [-np.sin(h), np.cos(h)]])
    b = torch.Tensor([crop[2] - crop[0], crop[3] - crop[1]]) / 2
    b = A.matmul(-b) + b
    rot2, tran2 = A.matmul(rot2), A.matmul(tran2) + b
    rot, tran = torch.eye(3), torch.zeros(3)
    rot[:2, :2], tran[:2] = rot2, tran2
    return rot, tran


def camera_to_lss(camera_to_lidar):
    """UE4 camera axes -> the Lift-Splat convention (`ue4_to_lss`, utils/camera_utils.py:553-568)"""
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.array([[0, 0, 1], [1, 0, 0], [0, -1, 0]], dtype=np.float32)
    return np.matmul(np.linalg.inv(camera_to_lidar), T)


_IMG_MEAN = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
_IMG_STD = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)


def _augment_pil(img, resize_dims, crop, flip, rotate):
    """resize, crop, flip, rotate of `img_transform` (utils/camera_utils.py:64-72), PIL's default resampling per mode"""
    from PIL import Image
    img = img.resize(resize_dims).crop(crop)
    if flip:
        img = img.transpose(method=Image.FLIP_LEFT_RIGHT)
    return img.rotate(rotate)


def _image_tensor(img, resize_dims, crop, flip, rotate):
    """PIL image -> normalised [3,H,W] tensor (`normalize_img` :136-143 = ToTensor + ImageNet mean / std)"""
    img = _augment_pil(img, resize_dims, crop, flip, rotate)
    a = np.array(img.convert("RGB") if img.mode != "RGB" else img, dtype=np.uint8)       # a writable copy
    t = torch.from_numpy(a).permute(2, 0, 1).to(torch.float32).div(255)
    return (t - _IMG_MEAN) / _IMG_STD


def _depth_tensor(depth_img, resize_dims, crop, flip, rotate):
    """CARLA depth PNG (24-bit value in R + 256 G + 65536 B, full scale 1000 m) -> 16-bit image -> the colour image's
    augmentation -> [1,H,W] metres (`decode_depth_carla` :145-166, `pil_depth_to_tensor` :192-210)"""
    from PIL import Image
    d = np.array(depth_img).astype(np.uint32)
    d = (d[:, :, 0] + d[:, :, 1] * 256 + d[:, :, 2] * 256 * 256).astype(np.float64) / (256 * 256 * 256 - 1) * 1000
    pil = Image.fromarray(np.clip(d * 65535 / 1000, 0, 65535).astype(np.uint16))      # uint16 -> mode "I;16"
    pil = _augment_pil(pil, resize_dims, crop, flip, rotate)
    if pil.mode != "I;16":
        raise ValueError("depth image lost its 16-bit mode in the augmentation")
    return torch.from_numpy(np.array(pil, dtype=np.float32) * 1000 / 65535.0).unsqueeze(0)


# ----------------------------------------------------------------------------------------------------- the dataset
class IntermediateFusionDatasetAirv2x(torch.utils.data.Dataset):
    """`IntermediateFusionDatasetAirv2x(params, visualize, train, source=...)`: same constructor keys, `__getitem__`
    bookkeeping and `collate_batch_train` layout as the reference class, with the lidar payload as `raw_points` and the
    labels as padded boxes (see the module docstring). `source[idx]` (or `source(idx)`) returns what `retrieve_base_data`
    returns: `base_data_dict` or `(base_data_dict, scenario_index, timestamp_key)`.

    shuffle: True = `np.random.permutation` per agent in the reference's order (`shuffle_points`, pcd_utils.py:193-197);
    False = keep the point order (first-come voxelisation depends on it)."""

    def __init__(self, params, visualize=False, train=True, source=None, shuffle=True, pin_memory=False):
        self.params, self.visualize, self.train, self.training = params, visualize, train, train
        if source is None:      # the reference's behaviour: scan params["root_dir"] / ["validate_dir"] (basedataset.py:73-207)
            root = params.get("root_dir" if train else "validate_dir")
            if root and os.path.isdir(root):
                from .airv2x_scenes import AirV2XScenes
                source = AirV2XScenes(params, train)
        self.source, self.shuffle, self.pin_memory = source, shuffle, pin_memory
        fa = params["fusion"]["args"]
        assert "proj_first" in fa
        self.proj_first = bool(fa["proj_first"])
        self.cur_ego_pose_flag = fa.get("cur_ego_pose_flag", True)
        self.collaborators = list(params["collaborators"])
        self.active_sensors = list(params["active_sensors"])
        self.use_cam, self.use_lidar = "cam" in self.active_sensors, "lidar" in self.active_sensors
        mc = params.get("train_params", {}).get("max_cav")
        self.max_cav = dict(mc) if mc else {"vehicle": 10, "rsu": 5, "drone": 5}          # basedataset.py:141-149
        self.max_cav_num = sum(self.max_cav[c] for c in self.collaborators)
        self.aug_conf = {"vehicle": fa.get("veh_data_aug_conf"), "rsu": fa.get("rsu_data_aug_conf"),
                         "drone": fa.get("drone_data_aug_conf")}
        self.ego_type = params.get("ego_type", "vehicle")
        assert self.ego_type in MODEL_ORDER, "ego type %s not supported" % self.ego_type
        self.agent_order = [self.ego_type] + [t for t in MODEL_ORDER if t != self.ego_type]   # :129-134
        pp = params["postprocess"]
        self.max_num, self.order = int(pp["max_num"]), pp["order"]
        self.gt_range = pp["anchor_args"]["cav_lidar_range"]
        self._post = None

    # -- source ------------------------------------------------------------------------------------------
    def __len__(self):
        if self.source is None or not hasattr(self.source, "__len__"):
            raise TypeError("this dataset has no sized scene source")
        return len(self.source)

    def retrieve_base_data(self, idx):
        if self.source is None:
            raise NotImplementedError("no scene source: params['root_dir'] / ['validate_dir'] is not a directory and no source= "
                                      "(a sequence or callable yielding what basedataset.retrieve_base_data returns) was given")
        rec = self.source(idx) if callable(self.source) else self.source[idx]
        if isinstance(rec, tuple):
            return rec
        return rec, 0, idx

    def __getitem__(self, idx):
        base, scenario_index, timestamp_key = self.retrieve_base_data(idx)
        return self.assemble(base, scenario_index, timestamp_key)

    # -- one agent -----------------------------------------------------------------------------------------
    def _cam_inputs(self, rec):
        """camera tensors + geometry of one agent (:497-580); draws the augmentation in the reference's order"""
        cams = rec.get("cameras") or []
        if not cams:
            return None
        depth = rec.get("depth") or []           # optional 4th image channel: metric depth for the depth-supervised lift
        n = len(cams)
        ext = np.asarray(rec["params"]["delay_extrinsic"]).reshape(n, 4, 4)
        intr = np.asarray(rec["params"]["delay_intrinsic"]).reshape(n, 3, 3)
        out = {k: [] for k in ("imgs", "intrinsics", "extrinsics", "rots", "trans", "post_rots", "post_trans")}
        for i, img in enumerate(cams):
            c2l = camera_to_lss(ext[i])
            resize, dims, crop, flip, rotate = sample_augmentation(self.aug_conf[rec["agent_type"]], self.train)
            rot, tran = post_homography(resize, crop, flip, rotate)
            planes = [_image_tensor(img, dims, crop, flip, rotate)]
            if depth:
                planes.append(_depth_tensor(depth[i], dims, crop, flip, rotate))
            out["imgs"].append(torch.cat(planes, dim=0))
            out["intrinsics"].append(torch.from_numpy(intr[i]))
            out["extrinsics"].append(torch.from_numpy(c2l))
            out["rots"].append(torch.from_numpy(c2l[:3, :3]))
            out["trans"].append(torch.from_numpy(c2l[:3, 3]))
            out["post_rots"].append(rot)
            out["post_trans"].append(tran)
        return {k: torch.stack(v) for k, v in out.items()}

    def get_item_single_car(self, rec, ego_pose, cache=None):
        """ground truth of this agent's surroundings in the ego frame + its raw cloud (:456-618). The cloud is NOT filtered
        or projected here: it is handed to the GPU with `transformation_matrix`."""
        boxes, ids, classes = project_world_objects(rec["params"]["objects"], ego_pose, self.gt_range, self.order, cache)
        if len(ids) > self.max_num:     # the reference writes into a [max_num, 7] array (base_postprocessor.py:614-621)
            raise IndexError("more than max_num = %d objects around one agent" % self.max_num)
        cam = self._cam_inputs(rec)
        cloud = np.asarray(rec["lidar_np"], dtype=np.float32).reshape(-1, 4)
        if self.shuffle:
            cloud = cloud[np.random.permutation(cloud.shape[0])]
        return {"agent_type": rec["agent_type"], "cam_inputs": cam, "lidar_np": cloud,
                "transformation_matrix": np.asarray(rec["params"]["transformation_matrix"], dtype=np.float64),
                "object_bbx_center": boxes, "object_ids": ids, "class_ids": classes}

    # -- one scene -----------------------------------------------------------------------------------------
    def get_pairwise_transformation(self, base_data_dict, max_cav, cop_agent_type):
        """([L,L,4,4] lidar, [L,L,4,4] image) agent i -> agent j transforms (:967-1020): identity everywhere for the lidar
        features under `proj_first`, `inv(T_j) T_i` (zero beyond the present agents) otherwise and for the images."""
        pair = np.zeros((max_cav, max_cav, 4, 4))
        ts = [np.asarray(c["params"]["transformation_matrix"]) for c in base_data_dict.values()
              if c["agent_type"] in cop_agent_type]
        inv = [np.linalg.inv(t) for t in ts]
        for i, ti in enumerate(ts):
            for j in range(len(ts)):
                pair[i, j] = np.eye(4) if i == j else np.dot(inv[j], ti)
        if self.proj_first:
            ident = np.zeros((max_cav, max_cav, 4, 4))
            ident[:, :] = np.identity(4)
            return ident, pair
        return pair, pair

    def assemble(self, base_data_dict, scenario_index=0, timestamp_key=0):
        """`__getitem__` after `retrieve_base_data` (:137-422): prune by communication range, keep the `max_cav` nearest
        agents per type, order the types ego-type first, merge the ground truth (first occurrence of an object id wins),
        pad the per-agent priors to `max_cav_num`."""
        ego_id, ego_pose = -1, []
        for cid, rec in base_data_dict.items():
            if rec["ego"]:
                ego_id, ego_pose = cid, rec["params"]["delay_ego_lidar_pose"]
                break
        per_type = {t: [] for t in MODEL_ORDER}
        ego_rec, projected = None, {}           # object records shared between agents are projected once per scene
        for cid, rec in base_data_dict.items():
            t = rec["agent_type"]
            if t not in self.collaborators:     # not a collaborator of this yaml: the models hold no encoder for it (the
                continue                        # reference counts such an agent and then mis-pads / fails in np.tile)
            if rec["distance_to_ego"] > COM_RANGE[t]:
                continue
            item = self.get_item_single_car(rec, ego_pose, projected)
            item.update(cav_id=cid, distance=rec["distance_to_ego"],
                        velocity=rec["params"]["odometry"]["ego_speed"] / 30.0, time_delay=float(rec["time_delay"]),
                        spatial_correction_matrix=np.asarray(rec["params"]["spatial_correction_matrix"]))
            per_type[t].append(item)
            if rec["ego"]:
                ego_rec = rec
        kept = OrderedDict()
        boxes, obj_ids, cls_ids, velocity, delay, infra, corr = [], [], [], [], [], [], []
        for t in self.agent_order:
            if not per_type[t]:
                continue
            idx, _ = nearest_agents([(a["cav_id"], a["distance"]) for a in per_type[t]], self.max_cav[t])
            per_type[t] = [per_type[t][i] for i in idx]
            for a in per_type[t]:
                kept.setdefault(a["cav_id"], base_data_dict[a["cav_id"]])
                # an agent whose cloud is empty supervises nothing (:603-604); its ids are dropped with its boxes here
                # (the reference keeps the ids and then indexes past its box stack)
                if a["lidar_np"].shape[0] > 0:
                    boxes.append(a["object_bbx_center"])
                    obj_ids.extend(a["object_ids"])
                    cls_ids.extend(a["class_ids"])
                velocity.append(a["velocity"])
                delay.append(a["time_delay"])
                infra.append(INFRA[t])
                corr.append(a["spatial_correction_matrix"])
        n_total = sum(len(v) for v in per_type.values())
        pair, img_pair = self.get_pairwise_transformation(kept, self.max_cav_num, self.collaborators)
        pad = self.max_cav_num - n_total
        corr = np.concatenate([np.stack(corr), np.tile(np.eye(4)[None], (pad, 1, 1))], axis=0)
        first = {}
        for i, oid in enumerate(obj_ids):
            first.setdefault(oid, i)
        sel = list(first.values())
        stack = np.vstack(boxes)[sel] if boxes and len(obj_ids) else np.zeros((0, 7))
        if stack.shape[0] > self.max_num:
            raise ValueError("%d objects in the scene, postprocess.max_num is %d" % (stack.shape[0], self.max_num))
        gt = np.zeros((self.max_num, 7))
        mask = np.zeros(self.max_num)
        cls_pad = np.zeros(self.max_num, dtype=np.int64)
        gt[:len(sel)], mask[:len(sel)] = stack, 1
        cls_pad[:len(sel)] = np.asarray([cls_ids[i] for i in sel], dtype=np.int64)
        ego = {"ego_id": ego_id, "object_bbx_center": gt, "object_bbx_mask": mask, "object_class_ids": cls_pad,
               "object_ids": [obj_ids[i] for i in sel], "class_ids": [cls_ids[i] for i in sel],
               "num_cavs": n_total, "pairwise_t_matrix_collab": pair, "img_pairwise_t_matrix_collab": img_pair,
               "spatial_correction_matrix": corr,
               "velocity": velocity + pad * [0.0], "time_delay": delay + pad * [0.0], "infra": infra + pad * [0.0],
               "scenario_index": scenario_index, "timestamp_key": timestamp_key,
               "metadata_path": None if ego_rec is None else ego_rec.get("metadata_path"), "ego_lidar_pose": ego_pose}
        if ego_rec is not None:
            for k in ("dynamic_seg_label", "static_seg_label"):
                if ego_rec.get(k) is not None:
                    ego[k] = ego_rec[k]
        for t in MODEL_ORDER:
            a = ABBR[t]
            ego["num_" + a] = len(per_type[t])
            ego["lidar_%s_list" % a] = [x["lidar_np"] for x in per_type[t]]
            ego["transformation_matrix_%s_list" % a] = [x["transformation_matrix"] for x in per_type[t]]
            ego["cav_ids_" + a] = [x["cav_id"] for x in per_type[t]]
            cams = [x["cam_inputs"] for x in per_type[t] if x["cam_inputs"] is not None]
            ego["merged_cam_inputs_dict_" + a] = ({k: torch.stack([c[k] for c in cams]) for k in cams[0]}
                                                  if cams else {})                      # merge "stack" (:316-318)
        return OrderedDict(ego=ego)

    # -- batch -----------------------------------------------------------------------------------------------
    def collate_batch_train(self, batch):
        """`collate_batch_train` (:620-888) with the B200 payload: `raw_points` (scene-major, vehicles / RSUs / drones per
        scene — the order the models repack to), padded boxes instead of label maps."""
        egos = [b["ego"] for b in batch]
        out = {"object_bbx_center": torch.from_numpy(np.array([e["object_bbx_center"] for e in egos])),
               "object_bbx_mask": torch.from_numpy(np.array([e["object_bbx_mask"] for e in egos])),
               "object_class_ids": torch.from_numpy(np.array([e["object_class_ids"] for e in egos])),
               "object_ids": [e["object_ids"] for e in egos], "class_ids": [e["class_ids"] for e in egos],
               "record_len": torch.from_numpy(np.array([e["num_cavs"] for e in egos], dtype=np.int32)),
               "pairwise_t_matrix_collab": torch.from_numpy(np.array([e["pairwise_t_matrix_collab"] for e in egos])).float(),
               "img_pairwise_t_matrix_collab":
                   torch.from_numpy(np.array([e["img_pairwise_t_matrix_collab"] for e in egos])).float(),
               "prior_encoding": torch.stack([torch.from_numpy(np.array([e[k] for e in egos]))
                                              for k in ("velocity", "time_delay", "infra")], dim=-1).float(),
               "spatial_correction_matrix": torch.from_numpy(np.array([e["spatial_correction_matrix"] for e in egos])),
               "scenario_index_list": [e["scenario_index"] for e in egos],
               "timestamp_key_list": [e["timestamp_key"] for e in egos],
               "metadata_path_list": [e["metadata_path"] for e in egos],
               "ego_lidar_pose_list": [e["ego_lidar_pose"] for e in egos]}
        # what `criterion(output_dict, batch["ego"]["label_dict"])` of tools/train.py:222-226 receives: the boxes, from which
        # this repo's criterion assigns the anchor targets on the GPU (det_loss._Criterion._targets)
        out["label_dict"] = {"object_bbx_center": out["object_bbx_center"], "object_bbx_mask": out["object_bbx_mask"],
                             "object_class_ids": out["object_class_ids"], "postprocess": self.params["postprocess"]}
        if all("dynamic_seg_label" in e and "static_seg_label" in e for e in egos):
            out["seg_label_dict"] = {k: torch.from_numpy(np.array([e[k] for e in egos]))
                                     for k in ("dynamic_seg_label", "static_seg_label")}
        for t in MODEL_ORDER:
            a = ABBR[t]
            present = [i for i, e in enumerate(egos) if e["num_" + a] > 0]
            cams = [egos[i]["merged_cam_inputs_dict_" + a] for i in present if egos[i]["merged_cam_inputs_dict_" + a]]
            out[t] = {"batch_merged_lidar_features_torch": None,      # voxelisation happens on the GPU: see raw_points
                      "batch_merged_cam_inputs": ({k: torch.cat([c[k] for c in cams], dim=0) for k in cams[0]}
                                                  if cams else {}),     # merge "cat" (:768-770)
                      "record_len": torch.from_numpy(np.array([e["num_" + a] for e in egos], dtype=np.int32)),
                      "batch_idxs": present}
        clouds, xforms = [], []
        for e in egos:
            for t in MODEL_ORDER:
                if t in self.collaborators:
                    clouds += e["lidar_%s_list" % ABBR[t]]
                    xforms += e["transformation_matrix_%s_list" % ABBR[t]]
        sizes = np.array([0] + [c.shape[0] for c in clouds], dtype=np.int64)
        points = torch.from_numpy(np.concatenate(clouds, axis=0) if clouds else np.zeros((0, 4), np.float32))
        if self.pin_memory:
            points = points.pin_memory()
        # proj_first False: the clouds stay in their own frames (identity "projection"); every agent's body box goes either way
        tf = np.stack(xforms) if self.proj_first else np.tile(np.eye(4)[None], (len(clouds), 1, 1))
        out["raw_points"] = {"points": points, "offsets": torch.from_numpy(np.cumsum(sizes).astype(np.int32)),
                             "preprocess": self.params["preprocess"], "filter": True,
                             "transforms": torch.from_numpy(tf.reshape(-1, 4, 4))}
        return {"ego": out}

    def collate_batch_test(self, batch):
        """:891-911: batch of one, plus the anchors and the identity `transformation_matrix` of the ego"""
        assert len(batch) <= 1, "Batch size 1 is required during testing!"
        out = self.collate_batch_train(batch)
        from .postprocess import generate_anchor_box
        pp = self.params["postprocess"]
        out["ego"]["anchor_box"] = torch.from_numpy(np.array(generate_anchor_box(pp["anchor_args"], pp["order"])))
        out["ego"]["transformation_matrix"] = torch.from_numpy(np.identity(4)).float()
        return out

    def generate_gt_bbx(self, data_dict):
        """ground truth of the evaluation (`generate_gt_bbx_airv2x`, data_utils/post_processor/base_postprocessor.py:118-205):
        the ego's valid boxes as [m,8,3] fp32 corners, restricted to the boxes whose 8 corners lie inside the x / y
        detection range, with their class labels and track ids. The corner arithmetic is the reference's tensor path
        (`boxes_to_corners_3d` on a float64 tensor: extents in fp64, the z-rotation as an fp32 matmul, centre added last)."""
        ego = data_dict["ego"] if "ego" in data_dict else data_dict
        box, mask = torch.as_tensor(ego["object_bbx_center"]), torch.as_tensor(ego["object_bbx_mask"])
        b = box[mask == 1].double().cpu()
        ids, classes = list(ego["object_ids"][0]), list(ego["class_ids"][0])
        if self.order == "hwl":
            b = b[:, [0, 1, 2, 5, 4, 3, 6]]
        half = b.new_tensor(_CORNER_SIGNS) / 2
        c = b[:, None, 3:6].repeat(1, 8, 1) * half[None]
        cos, sin = torch.cos(b[:, 6]), torch.sin(b[:, 6])
        zero, one = torch.zeros_like(cos), torch.ones_like(cos)
        rot = torch.stack((cos, sin, zero, -sin, cos, zero, zero, zero, one), dim=1).view(-1, 3, 3).float()
        corners = torch.matmul(c.float(), rot)
        corners += b[:, None, 0:3]
        first = {}
        for i, oid in enumerate(ids):
            first.setdefault(oid, i)
        sel = list(first.values())
        corners, classes, ids = corners[sel], [classes[i] for i in sel], [ids[i] for i in sel]
        lo = torch.Tensor(self.gt_range[:2]).reshape(1, 1, -1)
        hi = torch.Tensor(self.gt_range[3:5]).reshape(1, 1, -1)
        keep = torch.all(torch.all(corners[:, :, :2] >= lo, dim=-1) & torch.all(corners[:, :, :2] <= hi, dim=-1), dim=-1)
        idx = keep.nonzero(as_tuple=True)[0].tolist()
        return corners[keep], [classes[i] for i in idx], [ids[i] for i in idx]

    def post_process(self, data_dict, output_dict):
        """:913-938: (pred_box_tensor, pred_score, pred_labels, pred_boxes3d, gt_box_tensor, gt_class_labels, gt_track_ids);
        predictions = GPU decode + rotated NMS of the ego's output (`postprocess.DetPostprocessor`)"""
        from .postprocess import DetPostprocessor
        out = output_dict["ego"] if "ego" in output_dict else output_dict
        dev = out["psm"].device
        if self._post is None or self._post[0] != dev:
            self._post = (dev, DetPostprocessor(self.params["postprocess"], dev))
        pred = self._post[1](out)
        gt, classes, tracks = self.generate_gt_bbx(data_dict)
        return (*pred, gt.to(dev), classes, tracks)
