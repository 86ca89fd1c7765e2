"""CPU (-m "not gpu"): the point-projection restatement (oracle/voxelize.py) against the reference's own torch lines."""
import numpy as np


def test_projection_restatement_matches_reference_formula():
    """oracle.project_points == the reference's own lines (F.pad + torch.einsum in fp32, box_utils.py:1055-1061) and, for
    dataset-sized clouds, == the fused multiply-add chain the CUDA kernel evaluates (csrc/voxelize.cu vox_project)."""
    import math

    import torch
    import torch.nn.functional as F

    from oracle import voxelize as V

    g = np.random.default_rng(3)
    P = 5000
    pts = np.stack([g.normal(0, 30, P), g.normal(0, 14, P), g.uniform(-3, 1, P), g.uniform(0, 1, P)], 1).astype(np.float32)
    T = np.eye(4)
    a = 0.83
    T[:2, :2] = [[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]]
    T[:3, 3] = [12.345678, -7.654321, 0.3]
    T[2, 0], T[0, 2] = 0.01, -0.02
    got = V.project_points(pts, T)
    ref = torch.einsum("ik, jk->ij", F.pad(torch.from_numpy(pts[:, :3].copy()).float(), (0, 1), mode="constant", value=1),
                       torch.from_numpy(T).float())[:, :3].numpy()
    assert np.array_equal(got[:, :3].view(np.uint32), ref.view(np.uint32)) and np.array_equal(got[:, 3], pts[:, 3])
    p, t = pts.astype(np.float64), T.astype(np.float32).astype(np.float64)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)
    for j in range(3):       # double holds the exact product of two floats: one rounding per fused step
        s = f32(p[:, 0] * t[j, 0])
        s = f32(s + p[:, 1] * t[j, 1])
        s = f32(s + p[:, 2] * t[j, 2])
        s = f32(s + t[j, 3])
        assert np.array_equal(s.astype(np.float32), got[:, j]), j
    assert V.dataset_points(pts, T, [-140.8, -40, -3, 140.8, 40, 1]).shape[0] < P
