"""TEST INFRASTRUCTURE (oracle). PARITY UNPINNED vs real spconv (absent third-party dep; see voxelize.c header).

numpy-facing wrapper of oracle/voxelize.c (compiled with gcc into oracle/_build/) and a pure-Python twin for
small cases. Follows opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:59-72, :96-116 and the
collate at :142-175 (concatenate agents, prepend the agent index column).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_voxelize.so")
_lib = None


def build():
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    src = os.path.join(_HERE, "voxelize.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.a2x_oracle_voxelize.restype = ctypes.c_int
    return _lib


def grid_size(lidar_range, voxel_size):
    """(nx, ny, nz) = round((hi - lo) / vs)   (sp_voxel_preprocessor.py:52-55)"""
    r = np.asarray(lidar_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    return tuple(int(x) for x in np.round((r[3:6] - r[0:3]) / v).astype(np.int64))


def voxelize(points, lidar_range, voxel_size, max_points=32, max_voxels=32000, return_point_voxel=False):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, f = pts.shape
    rng = np.ascontiguousarray(lidar_range, dtype=np.float32)
    vs = np.ascontiguousarray(voxel_size, dtype=np.float32)
    voxels = np.zeros((max_voxels, max_points, f), dtype=np.float32)
    coords = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    pv = np.full((n,), -1, dtype=np.int32)
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int32)
    m = _load().a2x_oracle_voxelize(pts.ctypes.data_as(fp), n, f, rng.ctypes.data_as(fp), vs.ctypes.data_as(fp),
                                    max_points, max_voxels, voxels.ctypes.data_as(fp), coords.ctypes.data_as(ip),
                                    num.ctypes.data_as(ip), pv.ctypes.data_as(ip))
    if m < 0:
        raise MemoryError("oracle voxelize failed")
    out = {"voxel_features": voxels[:m].copy(), "voxel_coords": coords[:m].copy(), "voxel_num_points": num[:m].copy()}
    if return_point_voxel:
        out["point_voxel"] = pv
    return out


def voxelize_py(points, lidar_range, voxel_size, max_points=32, max_voxels=32000):
    """Pure-Python twin (small inputs only); same semantics, fp32 arithmetic through numpy scalars."""
    pts = np.asarray(points, dtype=np.float32)
    lo = np.asarray(lidar_range[:3], dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    grid = grid_size(lidar_range, voxel_size)
    table = {}
    voxels, coords, num = [], [], []
    for p in pts:
        c = []
        ok = True
        for j in range(3):
            q = math.floor(np.float32(np.float32(p[j] - lo[j]) / vs[j]))
            if q < 0 or q >= grid[j]:
                ok = False
                break
            c.append(int(q))
        if not ok:
            continue
        key = (c[2], c[1], c[0])
        v = table.get(key)
        if v is None:
            if len(voxels) >= max_voxels:
                continue
            v = len(voxels)
            table[key] = v
            voxels.append(np.zeros((max_points, pts.shape[1]), dtype=np.float32))
            coords.append(key)
            num.append(0)
        if num[v] < max_points:
            voxels[v][num[v]] = p
            num[v] += 1
    m = len(voxels)
    return {"voxel_features": np.stack(voxels) if m else np.zeros((0, max_points, pts.shape[1]), np.float32),
            "voxel_coords": np.asarray(coords, dtype=np.int32).reshape(m, 3),
            "voxel_num_points": np.asarray(num, dtype=np.int32)}


def collate(per_agent):
    """sp_voxel_preprocessor.py:142-175: concat agents and prepend the agent index to coords -> [M,4] (agent,z,y,x)."""
    feats = np.concatenate([d["voxel_features"] for d in per_agent], axis=0)
    nums = np.concatenate([d["voxel_num_points"] for d in per_agent], axis=0)
    coords = np.concatenate([np.pad(d["voxel_coords"], ((0, 0), (1, 0)), mode="constant", constant_values=i)
                             for i, d in enumerate(per_agent)], axis=0)
    return {"voxel_features": feats, "voxel_coords": coords.astype(np.int32), "voxel_num_points": nums}


def mask_points(points, lidar_range, ego_box=True):
    """utils/pcd_utils.py:136-190: drop points inside the ego box, keep points strictly inside the range."""
    p = np.asarray(points, dtype=np.float32)
    if ego_box:
        keep = ~((p[:, 0] >= -1.95) & (p[:, 0] <= 2.95) & (p[:, 1] >= -1.1) & (p[:, 1] <= 1.1))
        p = p[keep]
    r = lidar_range
    keep = (p[:, 0] > r[0]) & (p[:, 0] < r[3]) & (p[:, 1] > r[1]) & (p[:, 1] < r[4]) & (p[:, 2] > r[2]) & (p[:, 2] < r[5])
    return p[keep]


def project_points(points, transformation_matrix):
    """utils/box_utils.py:1038-1066 (project_points_by_matrix_torch) as the dataset calls it on `lidar_np[:, :3]`
    (data_utils/datasets/airv2x/intermediate_fusion_dataset.py:592-600): numpy inputs go through
    `torch.from_numpy(x).float()` (common_utils.py:36-39), the points are padded with a homogeneous 1 and contracted
    with the 4x4 matrix by torch.einsum in fp32. Restated with the same torch calls, so the fp32 evaluation order is
    torch's own (x*T0 then fused multiply-adds in k order for clouds of >= 100 points on this image's CPU sgemm).
    points: [P, 4] (x, y, z, intensity) -> [P, 4] fp32 with xyz projected."""
    import torch
    import torch.nn.functional as F

    p = np.asarray(points, dtype=np.float32).copy()
    if p.shape[0] == 0:
        return p
    xyz = torch.from_numpy(p[:, :3].copy()).float()
    T = torch.from_numpy(np.asarray(transformation_matrix)).float()
    hom = F.pad(xyz, (0, 1), mode="constant", value=1)
    p[:, :3] = torch.einsum("ik, jk->ij", hom, T)[:, :3].numpy()
    return p


def dataset_points(points, transformation_matrix, lidar_range):
    """get_item_single_car's cloud pipeline (intermediate_fusion_dataset.py:590-600) without the random shuffle:
    mask_ego_points (sensor frame) -> project to the ego frame -> mask_points_by_range."""
    p = np.asarray(points, dtype=np.float32)
    keep = ~((p[:, 0] >= -1.95) & (p[:, 0] <= 2.95) & (p[:, 1] >= -1.1) & (p[:, 1] <= 1.1))
    p = project_points(p[keep], transformation_matrix)
    return mask_points(p, lidar_range, ego_box=False)

