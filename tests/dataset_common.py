"""Synthetic `retrieve_base_data` records (what `basedataset.retrieve_base_data`, basedataset.py:217-303, hands to
`__getitem__`) for the dataset-side tests: seeded agents of the three types around an ego vehicle, world objects with ids
shared between agents, clouds in each agent's sensor frame, one camera per agent. TEST INFRASTRUCTURE."""
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _pose(rng, centre, spread, z):
    return [centre[0] + rng.uniform(-spread, spread), centre[1] + rng.uniform(-spread, spread), z,
            rng.uniform(-2, 2), rng.uniform(-180, 180), rng.uniform(-2, 2)]


def synth_scene(ds_mod, seed, n_veh=3, n_rsu=2, n_drone=2, n_obj=40, n_pts=3000, far=True, cameras=True,
                empty_cloud_agent=None, cur_ego_pose_flag=True, depth=False, obj_span=(130.0, 36.0), agent_spread=1.0, pts_sigma=(35.0, 15.0)):
    """-> OrderedDict cav_id -> record. `ds_mod` = the module under test (its pose helpers build the transforms; the
    reference's own `x1_to_x2` is compared with them separately)."""
    rng = np.random.default_rng(seed)
    ego_odom = [rng.uniform(-50, 50), rng.uniform(-50, 50), 0.3, 0.0, rng.uniform(-180, 180), 0.0]
    lidar_rel = [0.0, 0.0, 1.9, 0.0, 0.0, 0.0]
    agents = [("vehicle", True, ego_odom)]
    for i in range(n_veh - 1):
        agents.append(("vehicle", False, _pose(rng, ego_odom, 60 * agent_spread, 0.3)))
    for i in range(n_rsu):
        agents.append(("rsu", False, _pose(rng, ego_odom, 70 * agent_spread, 4.0)))
    for i in range(n_drone):
        agents.append(("drone", False, _pose(rng, ego_odom, 90 * agent_spread, 40.0)))
    if far:   # one vehicle beyond the 120 m communication range
        p = list(ego_odom)
        p[0] += 150.0
        agents.append(("vehicle", False, p))
    # interleave the types like a real scenario folder (agent ids are not grouped by type)
    order = [0] + list(rng.permutation(np.arange(1, len(agents))))
    ego_meta = {"lidar": {"lidar_pose": lidar_rel}, "odometry": {"ego_pos": ego_odom}}
    ego_delay_odom = list(ego_odom)
    if not cur_ego_pose_flag:
        ego_delay_odom[0] -= 1.5
        ego_delay_odom[4] += 3.0
    ego_delay_meta = {"lidar": {"lidar_pose": lidar_rel}, "odometry": {"ego_pos": ego_delay_odom}}
    ego_lidar = ds_mod.abs_world_pose(lidar_rel, ego_odom)
    ego_T = ds_mod.pose_to_matrix(ego_lidar)
    # world objects: boxes around the ego, a few far away; every agent lists a random subset (ids overlap)
    objects = {}
    for k in range(n_obj):
        r = 200.0 if k % 7 == 6 else 1.0
        local = np.array([rng.uniform(-obj_span[0], obj_span[0]) * r, rng.uniform(-obj_span[1], obj_span[1]) * r, rng.uniform(-2.2, -0.9), 1.0])
        w = ego_T @ local
        objects[100 + k] = {"location": [w[0], w[1], w[2], rng.uniform(-1, 1), rng.uniform(-180, 180), rng.uniform(-1, 1)],
                            "center": [rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.6, 0.9)],
                            "extent": [rng.uniform(1.6, 2.6), rng.uniform(0.7, 1.0), rng.uniform(0.7, 0.9)],
                            "class": int(rng.integers(1, 7))}
    base = OrderedDict()
    for slot, a in enumerate(order):
        t, is_ego, odom = agents[a]
        cid = 10 * slot + (0 if is_ego else 3)
        meta = {"lidar": {"lidar_pose": lidar_rel}, "odometry": {"ego_pos": odom}}
        pp = ds_mod.agent_pose_params(meta, meta, ego_meta, ego_delay_meta, cur_ego_pose_flag)
        mine = OrderedDict((k, v) for k, v in objects.items() if is_ego or rng.uniform() < 0.6)
        n = 0 if empty_cloud_agent == slot else n_pts
        cloud = np.stack([rng.normal(0, pts_sigma[0], n), rng.normal(0, pts_sigma[1], n), rng.uniform(-3, 1, n) - (odom[2] - 0.3),
                          rng.uniform(0, 1, n)], axis=1).astype(np.float32)
        cloud[: n // 50, :2] = rng.uniform(-1.0, 1.0, (n // 50, 2))          # points on the agent's own body
        rec = OrderedDict(ego=is_ego, agent_type=t, distance_to_ego=ds_mod.distance_to_ego(odom, ego_odom),
                          time_delay=0 if is_ego else int(rng.integers(0, 3)))
        params = dict(pp)
        params["odometry"] = {"ego_speed": float(rng.uniform(0, 25)), "ego_pos": odom}
        params["objects"] = mine
        if cameras:
            from PIL import Image
            img = (rng.uniform(0, 255, (72, 128, 3))).astype(np.uint8)
            rec["cameras"] = [Image.fromarray(img).resize((1280, 720))]
            yaw = np.radians(rng.uniform(-180, 180))
            ext = np.eye(4, dtype=np.float32)
            ext[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
            ext[:3, 3] = [1.2, 0.0, 0.4]
            params["delay_extrinsic"] = ext[None]
            params["delay_intrinsic"] = np.array([[[640.0, 0, 640.0], [0, 640.0, 360.0], [0, 0, 1.0]]], dtype=np.float32)
            rec["depth"] = ([Image.fromarray(rng.integers(0, 256, (72, 128, 3), dtype=np.uint8)).resize((1280, 720), Image.NEAREST)]
                            if depth else [])
        else:
            rec["cameras"], rec["depth"] = [], []
        rec["params"] = params
        rec["lidar_np"] = cloud
        rec["dynamic_seg_label"] = np.zeros((2, 2), dtype=np.int64) + slot
        rec["static_seg_label"] = np.ones((2, 2), dtype=np.int64) * slot
        rec["metadata_path"] = "scene%d/agent%d/meta.pkl" % (seed, cid)
        base[cid] = rec
    return base
