"""GPU (-m gpu): the V2X-ViT fusion kernels and the drop-in Airv2xV2XVit (BASELINE config 3) against the oracle (pinned
to the real reference) and the recorded golden vectors. Tolerance (north_star): logits max-abs <= 1e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import v2xvit_common as VC
from oracle import v2xvit_oracle as VO

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def ops():
    import a2x_import

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return a2x_import.pkg("ops")


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def test_hgt_fold_and_attention(ops):
    """folded typed projections + per-pixel multi-agent attention == HGTCavAttention (hmsa.py:117-158) up to to_out"""
    g = torch.Generator().manual_seed(0)
    n, H, W, C, heads, dh = 4, 5, 6, 256, 8, 32
    sd, pre = {}, "f"
    for name in ("q_linears", "k_linears", "v_linears", "a_linears"):
        for t in range(2):
            sd["%s.%s.%d.weight" % (pre, name, t)] = torch.randn(C, C, generator=g) * 0.05
            sd["%s.%s.%d.bias" % (pre, name, t)] = torch.randn(C, generator=g) * 0.1
    sd[pre + ".relation_att"] = torch.randn(4, heads, dh, dh, generator=g) * 0.2
    sd[pre + ".relation_msg"] = torch.randn(4, heads, dh, dh, generator=g) * 0.2
    x = torch.randn(1, n, H, W, C, generator=g)
    types = [0, 1, 1, 0]
    prior = torch.zeros(1, n, H, W, 3)
    prior[0, :, :, :, 2] = torch.tensor(types, dtype=torch.float32)[:, None, None]
    mask = (torch.rand(1, H, W, 1, n, generator=g) > 0.3).float()
    mask[..., 0] = 1.0
    # reference with identity output projection
    ident = {k: v for k, v in sd.items()}
    for t in range(2):
        ident["%s.a_linears.%d.weight" % (pre, t)] = torch.eye(C)
        ident["%s.a_linears.%d.bias" % (pre, t)] = torch.zeros(C)
    want = VO.hgt_attention(ident, pre, x, mask, prior, heads, dh)[0]
    cu = {k: v.cuda() for k, v in sd.items()}
    wf, bf = torch.empty(2, 5 * C, C, device="cuda"), torch.empty(2, 5 * C, device="cuda")
    pair = lambda nme, s: (cu["%s.%s.0.%s" % (pre, nme, s)], cu["%s.%s.1.%s" % (pre, nme, s)])
    ops.hgt_fold(pair("q_linears", "weight"), pair("q_linears", "bias"), pair("k_linears", "weight"), pair("k_linears", "bias"),
                 pair("v_linears", "weight"), pair("v_linears", "bias"), cu[pre + ".relation_att"], cu[pre + ".relation_msg"],
                 heads, wf, bf)
    xc = x[0].cuda()
    qkv = torch.stack([F.linear(xc[a], wf[types[a]], bf[types[a]]) for a in range(n)])     # fp32 projection (torch)
    out = ops.Act.empty((n, H, W, C), "cuda", True)
    km = mask[0, :, :, 0, :].permute(2, 0, 1).contiguous().cuda()
    ops.hgt_attention_fwd(qkv, torch.tensor(types, dtype=torch.int32).cuda(), km, heads, dh, out)
    assert rel(out.hi.cpu(), want) < 2e-5


def test_split_attn_and_rte(ops):
    g = torch.Generator().manual_seed(1)
    n, H, W, C = 3, 6, 8, 256
    wins = [torch.randn(1, n, H, W, C, generator=g) for _ in range(3)]
    sd = {"s.fc1.weight": torch.randn(C, C, generator=g) * 0.1, "s.bn1.weight": torch.rand(C, generator=g) + 0.5,
          "s.bn1.bias": torch.randn(C, generator=g) * 0.1, "s.fc2.weight": torch.randn(3 * C, C, generator=g) * 0.1}
    x = torch.randn(n, H, W, C, generator=g)
    want = x + VO.split_attn(sd, "s", wins)[0]
    xc = x.clone().cuda()
    cu = {k: v.cuda() for k, v in sd.items()}
    ops.split_attn_fuse(*[w[0].contiguous().cuda() for w in wins], cu["s.fc1.weight"], cu["s.bn1.weight"], cu["s.bn1.bias"],
                        cu["s.fc2.weight"], torch.empty(n, C, device="cuda"), torch.empty(n, 3, C, device="cuda"), xc)
    assert rel(xc.cpu(), want) < 1e-5
    # RTE
    emb = torch.randn(100, C, generator=g)
    lw, lb = torch.randn(C, C, generator=g) * 0.1, torch.randn(C, generator=g)
    idx = torch.tensor([0, 4, 2], dtype=torch.int32)
    want = x + F.linear(emb[idx.long()], lw, lb)[:, None, None, :]
    xc = x.clone().cuda()
    ops.rte_add(xc, emb.cuda(), idx.cuda(), lw.cuda(), lb.cuda(), torch.empty(n, C, device="cuda"))
    assert rel(xc.cpu(), want) < 1e-5


def test_sttf_warp_and_roi_mask(ops):
    """host geometry + warp kernel == STTF / get_roi_and_cav_mask of the reference (align_corners=True)"""
    import a2x_import

    Wp = a2x_import.pkg("warp")
    g = torch.Generator().manual_seed(2)
    agents = ["vehicle", "vehicle", "rsu", "drone"]
    _, scm = VC.scene_extras(agents, 4)
    H, W, C = 32, 64, 64
    x = torch.randn(1, 4, H, W, C, generator=g)
    want = VO.sttf(x, scm, 0.4, 4)[0]
    theta = Wp.sttf_theta(scm, 0.4, 4, H, W)[0].cuda()
    out = ops.Act(torch.empty(4, H, W, C, device="cuda"))
    ops.warp_affine_fwd(x[0].contiguous().cuda(), theta, out, align_corners=True)
    assert rel(out.hi[1:].cpu(), want[1:]) < 1e-5
    mask = VO.roi_and_cav_mask((1, 4, H, W, C), torch.tensor([[1, 1, 1, 0]]), scm, 0.4, 4)   # (1,H,W,1,L)
    km = torch.empty(4, H, W, device="cuda")
    ops.roi_mask(theta, torch.tensor([1, 1, 1, 0], dtype=torch.int32).cuda(), 4, H, W, km)
    want_m = mask[0, :, :, 0, :].permute(2, 0, 1)
    assert float((km.cpu() != want_m).float().mean()) < 2e-3          # nearest-neighbour ties on the ROI border
    assert float(km[3].abs().max()) == 0.0 and float(km[0].min()) == 1.0


@pytest.fixture(scope="module")
def small():
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_v2xvit")
    cfg, gold = VC.load_small()
    model = M.Airv2xV2XVit(cfg["model_args"])
    model.load_state_dict(VC.golden_state_dict(model, gold))
    model.cuda().eval()
    return cfg, gold, model


def test_eval_matches_reference_golden(small):
    import w2c_common as C

    cfg, gold, model = small
    dd = VC.golden_scene(cfg, gold)
    with torch.no_grad():
        out = model(C.to_device(dd, "cuda"))
    for k in ("psm", "rm", "obj"):
        assert out[k].shape == gold["eval_" + k].shape
        assert np.abs(out[k].cpu().numpy() - gold["eval_" + k]).max() < TOL, k
    assert out["comm_rate"] == int(gold["eval_comm_rate"])
