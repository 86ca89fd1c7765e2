"""Host-side geometry of the ego-warp (tiny 3x3 algebra per agent; the resampling itself is csrc/warp.cu).

Mirrors, in fp32 and in the reference's operation order so the sampling coordinates agree to rounding:
  get_discretized_transformation_matrix / get_rotation_matrix2d / get_transformation_matrix / normalize_homography /
  warp_affine (opencood/models/common_modules/torch_transformation_utils.py:116-143, :265-308, :203-262, :337-381)
  and the where2comm normalisation (opencood/models/where2comm_modules/where2comm_attn.py:293-307).
"""
import torch


def sttf_theta(scm, discrete_ratio, downsample_rate, H, W):
    """spatial_correction_matrix [B, L, 4, 4] -> theta [B, L, 2, 3] (fp32, CPU) for F.affine_grid(align_corners=True)
    semantics, i.e. what warp_affine samples with (STTF, v2xvit_basic.py:23-38; ROI mask, :15-53)."""
    m = scm.detach().cpu()
    B, L = m.shape[:2]
    m = m[:, :, [0, 1], :][:, :, :, [0, 1, 3]].clone()
    m[:, :, :, -1] = m[:, :, :, -1] / (discrete_ratio * downsample_rate)
    M = m.float().reshape(-1, 2, 3)
    n = M.shape[0]
    eye = torch.eye(3, dtype=torch.float32).repeat(n, 1, 1)
    shift, shift_inv, rot = eye.clone(), eye.clone(), eye.clone()
    center = torch.tensor([W / 2, H / 2], dtype=torch.float32)
    shift[:, :2, 2] = center
    shift_inv[:, :2, 2] = -center
    rot[:, :2, :2] = M[:, :2, :2]
    T = (shift @ rot @ shift_inv)[:, :2, :].clone()
    T[..., 2] += M[..., 2]
    M3 = torch.nn.functional.pad(T, [0, 0, 0, 1], "constant", 0.0).clone()
    M3[..., -1, -1] += 1.0
    norm = torch.tensor([[2.0 / (W - 1.0) if W > 1 else 2.0 / 1e-14, 0.0, -1.0],
                         [0.0, 2.0 / (H - 1.0) if H > 1 else 2.0 / 1e-14, -1.0], [0.0, 0.0, 1.0]], dtype=torch.float32)[None]
    dst_norm_trans_src_norm = norm @ (M3 @ torch.inverse(norm))
    return torch.inverse(dst_norm_trans_src_norm)[:, :2, :].reshape(B, L, 2, 3).contiguous()


def normalize_pairwise(pairwise_t_matrix, H, W, downsample_rate, discrete_ratio):
    """pairwise_t_matrix [B, L, L, 4, 4] -> [B, L, L, 2, 3] for warp_affine_simple (align_corners=False):
    where2comm_attn.py:293-307 / utils/transformation_utils.py:396-422"""
    t = pairwise_t_matrix.detach().cpu().float()[:, :, :, [0, 1], :][:, :, :, :, [0, 1, 3]].clone()
    t[..., 0, 1] = t[..., 0, 1] * H / W
    t[..., 1, 0] = t[..., 1, 0] * W / H
    t[..., 0, 2] = t[..., 0, 2] / (downsample_rate * discrete_ratio * W) * 2
    t[..., 1, 2] = t[..., 1, 2] / (downsample_rate * discrete_ratio * H) * 2
    return t.contiguous()
