"""GPU (-m gpu): the transformer-fusion token kernels and the drop-in Airv2xCoBEVT (BASELINE config 4) against the
oracle (pinned bit-exact to the real reference) and the recorded golden vectors.
Tolerance (north_star): logits max-abs <= 1e-3 (fp32 reference); per-kernel 1e-5 relative-to-max (fp32 kernels),
5e-5 for the bf16x3 tensor-core linears."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cobevt_common as CC
from oracle import cobevt_oracle as CO

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


def test_layernorm_and_agent_mean(ops):
    g = torch.Generator().manual_seed(0)
    for C in (128, 256, 384):
        x = (torch.randn(3, 5, 7, C, generator=g) * 3 + 1).cuda()
        gam, bet = torch.rand(C, generator=g).cuda() + 0.5, torch.randn(C, generator=g).cuda()
        out = ops.Act.empty(x.shape, "cuda", True)
        ops.layernorm_fwd(x, gam, bet, out)
        ref = F.layer_norm(x, (C,), gam, bet, 1e-5)
        assert rel(out.hi, ref) < 1e-5
        assert torch.equal(out.b16[0], out.hi.to(torch.bfloat16))
    B, L, C = 2, 3, 256
    x = torch.randn(B * L, 4, 6, C, generator=g).cuda()
    gam, bet = torch.rand(C, generator=g).cuda() + 0.5, torch.randn(C, generator=g).cuda()
    out = ops.Act.empty((B, 4, 6, C), "cuda", True)
    ops.agent_mean_layernorm(x, B, L, gam, bet, out)
    ref = F.layer_norm(x.view(B, L, 4, 6, C).mean(1), (C,), gam, bet, 1e-5)
    assert rel(out.hi, ref) < 1e-5


def test_regroup(ops):
    g = torch.Generator().manual_seed(1)
    src = torch.randn(5, 3, 4, 64, generator=g).cuda()
    start = torch.tensor([0, 2], dtype=torch.int32).cuda()
    length = torch.tensor([2, 3], dtype=torch.int32).cuda()
    out = ops.Act(torch.full((2 * 4, 3, 4, 64), 7.0, device="cuda"))
    ops.regroup(src, start, length, 2, 4, out)
    want, mask = CO.regroup(src.cpu().permute(0, 3, 1, 2), [2, 3], 4)
    assert torch.equal(out.hi.cpu().view(2, 4, 3, 4, 64), want.permute(0, 1, 3, 4, 2))
    assert mask.tolist() == [[1, 1, 0, 0], [1, 1, 1, 0]]


def _tc(case):
    """shapes the tcgen05 attention kernels take (csrc/window_attn_tc.cuh); bf16x3 operands -> 5e-5 instead of 2e-5"""
    B, L, H, W, heads, dh, w = case
    return w == 4 and dh == 32 and 32 <= L * 16 <= 128


@pytest.mark.parametrize("case", [(2, 7, 8, 16, 8, 32, 4), (1, 3, 8, 8, 4, 64, 2), (1, 5, 4, 12, 16, 16, 4),
                                  # tcgen05 path: 128 / 32 / 80-token windows (8, 2, 5 agents), more windows than CTAs
                                  (1, 8, 8, 8, 8, 32, 4), (2, 2, 8, 8, 4, 32, 4), (1, 5, 4, 12, 8, 32, 4), (1, 7, 100, 352, 8, 32, 4),
                                  # small windows (128 / n windows per CTA): the V2X-ViT pyramid shapes, partial groups
                                  (3, 1, 8, 12, 16, 16, 2), (2, 1, 8, 16, 8, 32, 4), (2, 1, 12, 8, 4, 64, 4),
                                  (2, 2, 4, 8, 8, 32, 2), (1, 1, 4, 12, 8, 32, 4)])
@pytest.mark.parametrize("grid_mode", [False, True])
def test_window_attention(ops, case, grid_mode):
    """softmax(q*scale k^T + rel-pos bias, masked keys) v per window / grid cell == Attention.forward without the
    linears (cobevt_modules/swap_fusion_modules.py:78-127)"""
    B, L, H, W, heads, dh, w = case
    D = heads * dh
    g = torch.Generator().manual_seed(2)
    qkv = torch.randn(B * L, H, W, 3 * D, generator=g)
    table = torch.randn((2 * L - 1) * (2 * w - 1) ** 2, heads, generator=g)
    mask = torch.ones(B, L, dtype=torch.int32)
    if L > 1:
        mask[0, L - 1] = 0           # trailing padded agents: the tcgen05 kernels never touch their key columns
        if L >= 5:
            mask[0, 2] = 0           # ... and a hole among the valid ones: masked by value
        if B > 1:
            mask[1, 1:] = 0
    out = ops.Act.empty((B * L, H, W, D), "cuda", True)
    ops.window_attention_fwd(qkv.cuda(), table.cuda(), mask.cuda(), B, L, heads, dh, w, grid_mode, out)
    # reference: identity linears around the oracle's attention core
    X, Y = H // w, W // w
    t = qkv.view(B, L, H, W, 3 * D)
    if grid_mode:
        t = t.view(B, L, w, X, w, Y, 3 * D).permute(0, 1, 3, 5, 2, 4, 6)
    else:
        t = t.view(B, L, X, w, Y, w, 3 * D).permute(0, 1, 2, 4, 3, 5, 6)
    tok = t.permute(0, 2, 3, 1, 4, 5, 6).reshape(B * X * Y, L * w * w, 3 * D)
    q, k, v = tok.chunk(3, -1)
    hs = lambda z: z.reshape(z.shape[0], z.shape[1], heads, dh).permute(0, 2, 1, 3)
    sim = (hs(q) * dh ** -0.5) @ hs(k).transpose(-1, -2)
    sim = sim + F.embedding(CO.relative_position_index(L, w), table).permute(2, 0, 1)
    km = mask.view(B, 1, 1, L, 1, 1).expand(B, X, Y, L, w, w).reshape(B * X * Y, 1, 1, L * w * w)
    sim = sim.masked_fill(km == 0, -float("inf"))
    o = (sim.softmax(-1) @ hs(v)).permute(0, 2, 1, 3).reshape(B, X, Y, L, w, w, D)
    if grid_mode:
        ref = o.permute(0, 3, 4, 1, 5, 2, 6).reshape(B * L, H, W, D)
    else:
        ref = o.permute(0, 3, 1, 4, 2, 5, 6).reshape(B * L, H, W, D)
    assert rel(out.hi.cpu(), ref) < (5e-5 if _tc(case) else 2e-5)
    assert torch.equal(out.b16[0], out.hi.to(torch.bfloat16))


def test_linear_gelu_residual(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 10, 256, generator=g).cuda()
    w1, b1 = (torch.randn(512, 256, generator=g) * 0.05).cuda(), torch.randn(512, generator=g).cuda() * 0.1
    w2, b2 = (torch.randn(256, 512, generator=g) * 0.05).cuda(), torch.randn(256, generator=g).cuda() * 0.1
    res = torch.randn(2, 6, 10, 256, generator=g).cuda()
    hid = ops.Act.empty((2, 6, 10, 512), "cuda", True)
    ops.linear_fwd(ops.split(x), ops.pack_conv_weight(w1.view(512, 256, 1, 1)), hid, bias=b1, act=2)
    want_h = F.gelu(F.linear(x, w1, b1))
    assert rel(hid.hi, want_h) < 5e-5
    y = ops.Act(res.clone())
    ops.linear_fwd(hid, ops.pack_conv_weight(w2.view(256, 512, 1, 1)), y, bias=b2, accumulate=True)
    assert rel(y.hi, res + F.linear(want_h, w2, b2)) < 5e-5


@pytest.mark.parametrize("align", [False, True])
def test_warp_affine(ops, align):
    """ego-warp == F.affine_grid + F.grid_sample (warp_affine_simple, torch_transformation_utils.py:327-334)"""
    g = torch.Generator().manual_seed(4)
    n, C, H, W = 3, 64, 20, 44
    x = torch.randn(n, C, H, W, generator=g).cuda().requires_grad_(True)
    ang = torch.tensor([0.0, 0.3, -2.1])
    theta = torch.stack([torch.stack([torch.cos(ang), -torch.sin(ang) * H / W, torch.tensor([0.0, 0.2, -0.4])], -1),
                         torch.stack([torch.sin(ang) * W / H, torch.cos(ang), torch.tensor([0.0, -0.1, 0.3])], -1)], 1).cuda()
    grid = F.affine_grid(theta, [n, C, H, W], align_corners=align)
    ref = F.grid_sample(x, grid, align_corners=align)
    dout = torch.randn(n, C, H, W, generator=g).cuda()
    ref.backward(dout)
    xn = x.detach().permute(0, 2, 3, 1).contiguous()
    out = ops.Act.empty((n, H, W, C), "cuda", True)
    ops.warp_affine_fwd(xn, theta, out, align_corners=align)
    assert rel(out.hi, ref.detach().permute(0, 2, 3, 1)) < 1e-5
    dsrc = torch.zeros_like(xn)
    ops.warp_affine_bwd(dout.permute(0, 2, 3, 1).contiguous(), theta, dsrc, align_corners=align)
    assert rel(dsrc, x.grad.permute(0, 2, 3, 1)) < 1e-5
    # identity transform (proj_first) is an exact copy
    eye = torch.tensor([[1.0, 0, 0], [0, 1.0, 0]]).repeat(n, 1, 1).cuda()
    ops.warp_affine_fwd(xn, eye, out, align_corners=align)
    assert rel(out.hi, xn) < 1e-5                                          # coordinates carry fp32 rounding
    nn_ref = F.grid_sample(x.detach(), grid, mode="nearest", align_corners=align)
    ops.warp_affine_fwd(xn, theta, out, align_corners=align, nearest=True)
    assert float((out.hi != nn_ref.permute(0, 2, 3, 1)).float().mean()) < 1e-3   # ties at .5 may round differently


@pytest.fixture(scope="module")
def small():
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    cfg, gold = CC.load_small()
    model = M.Airv2xCoBEVT(cfg["model_args"])
    sd = CC.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().eval()
    return cfg, gold, model, sd


def test_fusion_net_matches_reference_golden(small):
    """SwapFusionEncoder alone on the seeded ragged input recorded from the real reference"""
    cfg, gold, model, sd = small
    g = torch.Generator().manual_seed(int(gold["fusion_seed"]))
    x = torch.randn(2, 7, 256, 8, 16, generator=g)
    mask = torch.tensor([[1, 1, 1, 0, 0, 0, 0], [1, 1, 1, 1, 1, 0, 0]])
    x = x * mask[:, :, None, None, None]
    feat = torch.cat([x[0, :3], x[1, :5]]).permute(0, 2, 3, 1).contiguous().cuda()     # dense scene-major maps, NHWC
    eng = model.engine
    P = model._param_dict()
    eng._begin_step()
    W = eng._pack_weights(P)
    layout = {"record_len": [3, 5], "scene_start": torch.tensor([0, 3], dtype=torch.int32).cuda(),
              "scene_len": torch.tensor([3, 5], dtype=torch.int32).cuda(), "key_mask": mask.to(torch.int32).cuda()}
    fused = eng.fusion(P, W, feat, layout)
    got = fused.hi.permute(0, 3, 1, 2).cpu().numpy()
    assert np.abs(got - gold["fusion_out"]).max() < TOL


def test_eval_matches_reference_golden(small):
    import w2c_common as C

    cfg, gold, model, sd = small
    dd = CC.golden_scene(cfg, gold)
    with torch.no_grad():
        out = model(C.to_device(dd, "cuda"))
    for k in ("psm", "rm", "obj"):
        assert out[k].shape == gold["eval_" + k].shape
        assert np.abs(out[k].cpu().numpy() - gold["eval_" + k]).max() < TOL, k
    # train mode with grad enabled is the autograd-bridged path (torch.library ops; nn.Dropout(0.1) on, like the reference)
    model.train()
    tr = model(C.to_device(dd, "cuda"))
    assert tr["psm"].requires_grad and tr["psm"].shape == out["psm"].shape
    tr["psm"].sum().backward()
    assert all(p.grad is not None for p in model.parameters() if p.requires_grad)
    model.zero_grad()
    model.load_state_dict(sd)        # the train-mode forward moved the BatchNorm running statistics of the shared fixture
    model.eval()


def test_naive_compressor_eval_matches_reference_golden():
    """a10: NaiveCompressor (3 x conv3x3 + BN + ReLU, bias folded into the BN shift) between shrink and fusion"""
    import json

    import a2x_import
    import w2c_common as C

    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    cfg, gold = CC.load_small()
    args = json.loads(json.dumps(cfg["model_args"]))
    args["compression"] = 2
    model = M.Airv2xCoBEVT(args)
    model.load_state_dict(CC.golden_state_dict_compressed(model, gold))
    model.cuda().eval()
    with torch.no_grad():
        out = model(C.to_device(CC.golden_scene(cfg, gold), "cuda"))
    for k in ("psm", "rm", "obj"):
        assert np.abs(out[k].cpu().numpy() - gold["cmp2_eval_" + k]).max() < TOL, k


def test_layernorm_gelu_backward(ops):
    g = torch.Generator().manual_seed(20)
    x = (torch.randn(2, 5, 7, 256, generator=g) * 2 + 0.5).cuda().requires_grad_(True)
    gam = (torch.rand(256, generator=g) + 0.5).cuda().requires_grad_(True)
    bet = torch.randn(256, generator=g).cuda().requires_grad_(True)
    dy = torch.randn(2, 5, 7, 256, generator=g).cuda()
    F.layer_norm(x, (256,), gam, bet, 1e-5).backward(dy)
    acc = torch.randn(2, 5, 7, 256, generator=g).cuda()
    dx = acc.clone()
    dg, db = torch.zeros(256, dtype=torch.float64, device="cuda"), torch.zeros(256, dtype=torch.float64, device="cuda")
    ops.layernorm_bwd(x.detach(), dy, gam.detach(), dx, dg, db)
    assert rel(dx - acc, x.grad) < 1e-5 and rel(dg.float(), gam.grad) < 1e-5 and rel(db.float(), bet.grad) < 1e-5
    h = torch.randn(2, 5, 7, 256, generator=g).cuda().requires_grad_(True)
    F.gelu(h).backward(dy)
    y = ops.Act.empty(h.shape, "cuda", True)
    ops.gelu_fwd(h.detach(), y)
    d = ops.Act.empty(h.shape, "cuda", True)
    ops.gelu_bwd(dy, h.detach(), d)
    assert rel(y.hi, F.gelu(h.detach())) < 1e-6 and rel(d.hi, h.grad) < 1e-5


@pytest.mark.parametrize("case", [(2, 7, 8, 8, 8, 32, 4), (1, 3, 4, 8, 4, 16, 2),
                                  (1, 8, 8, 8, 8, 32, 4), (2, 2, 8, 8, 4, 32, 4), (1, 5, 4, 12, 8, 32, 4), (1, 7, 48, 64, 8, 32, 4),
                                  # small windows (V2X-ViT pyramid): 128 / n windows per CTA
                                  (3, 1, 8, 12, 16, 16, 2), (2, 1, 8, 16, 8, 32, 4), (2, 1, 12, 8, 4, 64, 4), (1, 1, 4, 12, 8, 32, 4),
                                  # 16 x 16 windows of the legacy V2X-ViT pyramid (256 tokens: recompute-from-rows kernel), 3 agents x 64
                                  (2, 1, 16, 32, 4, 64, 16), (1, 3, 8, 16, 2, 32, 8)])
@pytest.mark.parametrize("grid_mode", [False, True])
def test_window_attention_backward(ops, case, grid_mode):
    """dq, dk, dv and the relative-position-bias gradient == torch autograd of the attention core"""
    B, L, H, W, heads, dh, w = case
    D = heads * dh
    g = torch.Generator().manual_seed(21)
    qkv = torch.randn(B * L, H, W, 3 * D, generator=g, dtype=torch.float64).requires_grad_(True)
    table = torch.randn((2 * L - 1) * (2 * w - 1) ** 2, heads, generator=g, dtype=torch.float64).requires_grad_(True)
    mask = torch.ones(B, L, dtype=torch.int32)
    if L > 1:
        mask[0, L - 1] = 0
        if L >= 5:
            mask[0, 2] = 0
        if B > 1:
            mask[1, 1:] = 0
    dout = torch.randn(B * L, H, W, D, generator=g, dtype=torch.float64)
    X, Y = H // w, W // w
    t = qkv.view(B, L, H, W, 3 * D)
    t = t.view(B, L, w, X, w, Y, 3 * D).permute(0, 1, 3, 5, 2, 4, 6) if grid_mode else t.view(B, L, X, w, Y, w, 3 * D).permute(0, 1, 2, 4, 3, 5, 6)
    tok = t.permute(0, 2, 3, 1, 4, 5, 6).reshape(B * X * Y, L * w * w, 3 * D)
    q, k, v = tok.chunk(3, -1)
    hs = lambda z: z.reshape(z.shape[0], z.shape[1], heads, dh).permute(0, 2, 1, 3)
    sim = (hs(q) * dh ** -0.5) @ hs(k).transpose(-1, -2) + F.embedding(CO.relative_position_index(L, w), table).permute(2, 0, 1)
    km = mask.view(B, 1, 1, L, 1, 1).expand(B, X, Y, L, w, w).reshape(B * X * Y, 1, 1, L * w * w)
    o = (sim.masked_fill(km == 0, -float("inf")).softmax(-1) @ hs(v)).permute(0, 2, 1, 3).reshape(B, X, Y, L, w, w, D)
    ref = o.permute(0, 3, 4, 1, 5, 2, 6).reshape(B * L, H, W, D) if grid_mode else o.permute(0, 3, 1, 4, 2, 5, 6).reshape(B * L, H, W, D)
    ref.backward(dout)
    dqkv = torch.full((B * L, H, W, 3 * D), 9.0, device="cuda")
    dbias = torch.zeros(table.shape, device="cuda")
    ops.window_attention_bwd(qkv.detach().float().cuda(), dout.float().cuda(), table.detach().float().cuda(), mask.cuda(), B, L,
                             heads, dh, w, grid_mode, dqkv, dbias)
    tol = 5e-5 if _tc(case) else 2e-5
    assert rel(dqkv.cpu().double(), qkv.grad) < tol
    assert rel(dbias.cpu().double(), table.grad) < tol


def test_train_step_matches_oracle_autograd():
    """CoBEVT training step (train-mode BatchNorm, dropout off): loss and every parameter gradient against torch autograd
    through the oracle (itself pinned bit-exact to the real reference in eval mode). Gradients of the fusion network and
    heads are tight; encoder gradients pass ReLU / max gates (see tests/test_gpu_model.py) -> norm-wise bounds."""
    import json

    import a2x_import
    import w2c_common as C
    from oracle import w2c_oracle as O

    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    cfg, gold = CC.load_small()
    args = json.loads(json.dumps(cfg["model_args"]))
    args["fax_fusion"]["drop_out"] = 0.0
    model = M.Airv2xCoBEVT(args)
    sd = CC.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().train()
    dd = CC.golden_scene(cfg, gold)
    H, W = gold["eval_psm"].shape[2:]
    labels = O.make_labels(5, 1, H, W, args["anchor_number"])
    loss3 = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])
    # oracle: train-mode forward + the reference's loss + autograd on the CPU
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
         for k, v in sd.items()}
    torch.set_num_threads(8)
    out, _ = CO.cobevt_forward(p, args, dd, training=True)
    loss = O.point_pillar_loss_multiclass(out, labels, args["num_class"], cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])[0]
    loss.backward()
    assert abs(float(loss3.sum()) - float(loss.detach())) < 1e-3 * abs(float(loss.detach()))
    errs = {}
    for n, q in model.named_parameters():
        ref = p[n].grad
        if ref is None:
            continue
        errs[n] = float((q.grad.cpu() - ref).norm() / (ref.norm() + 1e-30))
    fusion = {n: e for n, e in errs.items() if n.startswith("fusion_net") or "head" in n}
    assert len(fusion) > 60 and max(fusion.values()) < 2e-2, sorted(fusion.items(), key=lambda kv: -kv[1])[:5]
    assert float(np.median(list(fusion.values()))) < 2e-3
    assert max(errs.values()) < 0.15 and float(np.median(list(errs.values()))) < 0.05, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    # reference-style use (tools/train.py:216-221): model(batch) -> the reference's loss (torch ops) -> loss.backward()
    g_step = {n: q.grad.clone() for n, q in model.named_parameters()}
    model.zero_grad()
    model.dropout = "off"
    out2 = model(C.to_device(dd, "cuda"))
    cpu_out = {k: out2[k].cpu() for k in ("psm", "rm", "obj")}
    loss2 = O.point_pillar_loss_multiclass(cpu_out, labels, args["num_class"], cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])[0]
    assert abs(float(loss2.detach()) - float(loss.detach())) < 1e-3 * abs(float(loss.detach()))
    loss2.backward()
    for n, q in model.named_parameters():
        assert q.grad is not None, n
        assert float((q.grad - g_step[n]).abs().max()) <= 2e-3 * float(g_step[n].abs().max()) + 1e-7, n


def test_train_step_with_naive_compressor_matches_oracle_autograd():
    """a10 in training: `compression: 2` puts NaiveCompressor (3 x conv3x3 + bias + BatchNorm(batch statistics) + ReLU,
    common_modules/naive_compress.py:10-42; its train mode is pinned to the real module by
    scripts/make_golden_compressor_train.py) between the shrink header and the fusion. Loss, the compressor's parameter
    gradients and its running statistics against torch autograd through the oracle; the conv biases are absorbed by the
    batch mean (gradient exactly zero here, ~1e-9 of rounding in autograd)."""
    import json

    import a2x_import
    import w2c_common as C
    from oracle import w2c_oracle as O

    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    cfg, gold = CC.load_small()
    args = json.loads(json.dumps(cfg["model_args"]))
    args["fax_fusion"]["drop_out"] = 0.0
    args["compression"] = 2
    model = M.Airv2xCoBEVT(args)
    sd = CC.golden_state_dict_compressed(model, gold)
    model.load_state_dict(sd)
    model.cuda().train()
    dd = CC.golden_scene(cfg, gold)
    H, W = gold["eval_psm"].shape[2:]
    labels = O.make_labels(5, 1, H, W, args["anchor_number"])
    loss3 = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
         for k, v in sd.items()}
    torch.set_num_threads(8)
    out, bufs = CO.cobevt_forward(p, args, dd, training=True)
    loss = O.point_pillar_loss_multiclass(out, labels, args["num_class"], cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])[0]
    loss.backward()
    assert abs(float(loss3.sum()) - float(loss.detach())) < 1e-3 * abs(float(loss.detach()))
    errs, scale = {}, 0.0
    for n, q in model.named_parameters():
        ref = p[n].grad
        if ref is None:
            continue
        if n.startswith("naive_compressor") and n.endswith(".bias") and n.split(".")[-2] in ("0", "3"):   # conv biases
            w_ref = p[n[:-len("bias")] + "weight"].grad
            assert float(q.grad.abs().max()) == 0.0 and float(ref.abs().max()) < 1e-5 * float(w_ref.abs().max()), n
            continue
        errs[n] = float((q.grad.cpu() - ref).norm() / (ref.norm() + 1e-30))
    cmp_ = {n: e for n, e in errs.items() if n.startswith("naive_compressor")}
    assert len(cmp_) == 9 and max(cmp_.values()) < 5e-2, sorted(cmp_.items(), key=lambda kv: -kv[1])
    fusion = {n: e for n, e in errs.items() if n.startswith("fusion_net") or "head" in n}
    assert max(fusion.values()) < 2e-2 and float(np.median(list(fusion.values()))) < 2e-3
    assert max(errs.values()) < 0.15 and float(np.median(list(errs.values()))) < 0.05, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    msd = model.state_dict()
    for k, v in bufs.items():
        if k.startswith("naive_compressor") and "num_batches" not in k:
            assert float((msd[k].cpu() - v).abs().max()) < 1e-5 + 1e-4 * float(v.abs().max()), k


def test_dropout_kernels_and_train_step_with_identical_masks(ops):
    """nn.Dropout of the fusion network (swap_fusion_modules.py:43, base_transformer.py:32,34) as counter-based masks:
    (i) the keep rate of the exported masks is 1 - p (binomial bounds) and sites / seeds give independent masks;
    (ii) dropout_apply / gelu_dropout fwd / bwd equal the torch expressions under the exported mask; (iii) the CoBEVT
    training step with the shipped drop_out = 0.1 equals torch autograd through the oracle fed the SAME masks."""
    import json

    import a2x_import
    import w2c_common as C
    from oracle import w2c_oracle as O

    # (i) statistics
    d = ops.Dropout(0.1, 12345)
    n = 1 << 20
    m0, m1 = ops.dropout_mask(n, d, 0).float(), ops.dropout_mask(n, d, 1).float()
    m2 = ops.dropout_mask(n, ops.Dropout(0.1, 12346), 0).float()
    sig = (0.1 * 0.9 / n) ** 0.5
    for m in (m0, m1, m2):
        assert abs(float(m.mean()) - 0.9) < 5 * sig
    assert abs(float((m0 * m1).mean()) - 0.81) < 1e-3 and abs(float((m0 * m2).mean()) - 0.81) < 1e-3   # independent
    assert abs(float((m0[:-1] * m0[1:]).mean()) - 0.81) < 1e-3                                          # no lag-1 correlation
    # (ii) kernels == torch under the exported mask
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 6, 8, 64, generator=g).cuda()
    r = torch.randn(4, 6, 8, 64, generator=g).cuda()
    d3 = ops.Dropout(0.3, 777)
    mk = ops.dropout_mask(x.numel(), d3, 5).view(x.shape).float()
    out = ops.Act.empty(x.shape, "cuda", True)
    ops.dropout_apply(x, d3, 5, out, residual=r)
    assert torch.allclose(out.hi, r + x * mk / 0.7, rtol=1e-6, atol=1e-6)
    assert torch.equal(out.b16[0], out.hi.to(torch.bfloat16))
    ops.gelu_dropout_fwd(x, d3, 5, out)
    assert torch.allclose(out.hi, F.gelu(x) * mk / 0.7, rtol=1e-5, atol=1e-6)
    xx = x.clone().requires_grad_(True)
    (F.gelu(xx) * mk / 0.7).backward(r)
    ops.gelu_dropout_bwd(r, x, d3, 5, out)
    assert torch.allclose(out.hi, xx.grad, rtol=1e-4, atol=1e-5)
    # (ii-b) x + dropout(linear(y) + bias) in the GEMM epilogue == the two-pass form under the exported mask, also when the
    # output is a slice of the site's tensor (elem_offset) and without residual
    tok = torch.randn(3, 10, 24, 128, generator=g).cuda()
    wl, bl = (torch.randn(64, 128, generator=g) * 0.1).cuda(), torch.randn(64, generator=g).cuda()
    res = torch.randn(3, 10, 24, 64, generator=g).cuda()
    pw = ops.pack_conv_weight(wl.view(64, 128, 1, 1))
    mk = ops.dropout_mask(res.numel(), d3, 9).view(res.shape).float()
    lin = ops.Act.empty(res.shape, "cuda", False)
    ops.linear_fwd(ops.split(tok), pw, lin, bias=bl)
    want = res + lin.hi * mk / 0.7
    y = torch.empty_like(res)
    ops.linear_dropout_residual_fwd(ops.split(tok), pw, y, bias=bl, residual=res, drop=d3, site=9)
    assert torch.allclose(y, want, rtol=1e-6, atol=1e-6)
    y2 = torch.zeros_like(res)
    ops.linear_dropout_residual_fwd(ops.split(tok).narrow_n(1, 2), pw, y2[1:3], bias=bl, residual=res[1:3], drop=d3, site=9,
                                    elem_offset=res[0].numel())
    assert torch.equal(y2[1:3], y[1:3]) and float(y2[0].abs().max()) == 0.0
    ops.linear_dropout_residual_fwd(ops.split(tok), pw, y2, bias=bl, drop=d3, site=9)
    assert torch.allclose(y2, lin.hi * mk / 0.7, rtol=1e-6, atol=1e-6)
    ops.linear_dropout_residual_fwd(ops.split(tok), pw, y2, bias=bl, residual=res)              # no dropout: plain residual
    assert torch.allclose(y2, res + lin.hi, rtol=1e-6, atol=1e-6)
    yy = res.clone()
    ops.linear_dropout_residual_fwd(ops.split(tok), pw, yy, bias=bl, residual=yy, drop=d3, site=9)   # residual aliases the output
    assert torch.equal(yy, y)
    # (iii) the training step with dropout ON
    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    cfg, gold = CC.load_small()
    args = json.loads(json.dumps(cfg["model_args"]))
    p_drop = float(args["fax_fusion"]["drop_out"])
    assert p_drop == 0.1                                   # the shipped yaml value
    model = M.Airv2xCoBEVT(args)
    sd = CC.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().train()
    dd = CC.golden_scene(cfg, gold)
    H, W = gold["eval_psm"].shape[2:]
    labels = O.make_labels(5, 1, H, W, args["anchor_number"])
    # train_step returns the engine's persistent [reg, cls, obj] buffer: clone to keep a value across steps
    loss3 = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"], dropout=4242).clone()
    loss_off = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"], dropout="off").clone()
    assert abs(float(loss3.sum()) - float(loss_off.sum())) > 1e-4 * abs(float(loss_off.sum()))      # dropout does something
    model.load_state_dict(sd)                                                                        # same BN running stats
    loss3b = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"], dropout=4242).clone()
    assert torch.allclose(loss3, loss3b, rtol=1e-6)                                                  # same seed, same step
    drop = model.last_dropout
    L, d_model, mlp = model.max_cav_num, args["fax_fusion"]["input_dim"], args["fax_fusion"]["mlp_dim"]
    n_tok = L * H * W
    widths = []
    for _ in range(args["fax_fusion"]["depth"]):
        for _ in range(2):
            widths += [d_model, mlp, d_model]              # to_out, gelu, net.3
    assert drop.n_sites == len(widths)
    masks = [ops.dropout_mask(n_tok * c, drop, s).cpu() for s, c in enumerate(widths)]
    pp = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd.items()}
    torch.set_num_threads(8)
    out_o, _ = CO.cobevt_forward(pp, args, dd, training=True, dropout=CO.MaskedDropout(p_drop, masks))
    loss = O.point_pillar_loss_multiclass(out_o, labels, args["num_class"], cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])[0]
    loss.backward()
    assert abs(float(loss3.sum()) - float(loss.detach())) < 1e-3 * abs(float(loss.detach()))
    errs = {}
    for nme, q in model.named_parameters():
        if pp[nme].grad is not None:
            errs[nme] = float((q.grad.cpu() - pp[nme].grad).norm() / (pp[nme].grad.norm() + 1e-30))
    fusion = {k: e for k, e in errs.items() if k.startswith("fusion_net") or "head" in k}
    assert len(fusion) > 60 and max(fusion.values()) < 2e-2, sorted(fusion.items(), key=lambda kv: -kv[1])[:5]
    assert float(np.median(list(fusion.values()))) < 2e-3


def test_ragged_multi_scene_batch_matches_oracle(small):
    """B = 3 ragged scenes (4, 2 and 3 agents): per-type collate, scene-major regrouping to L slots with the per-scene
    key mask, window / grid attention per scene — against the oracle on the CPU"""
    import w2c_common as C

    cfg, gold, model, sd = small
    model.eval()
    scenes = [["vehicle", "vehicle", "rsu", "drone"], ["vehicle", "drone"], ["vehicle", "rsu", "rsu"]]
    pre = dict(cfg["preprocess"])
    pre["args"] = dict(pre["args"])
    pre["args"]["max_voxel_test"] = pre["args"]["max_voxel_train"]
    dd, raw = C.make_batch(pre, scenes, 4000, 41, pre["args"]["max_voxel_train"])
    torch.set_num_threads(8)
    with torch.no_grad():
        ora, _ = CO.cobevt_forward(sd, cfg["model_args"], dd, training=False)
        out = model(C.to_device(dd, "cuda"))
        out_raw = model(raw)
    assert out["psm"].shape[0] == 3
    for k in ("psm", "rm", "obj"):
        assert float((out[k].cpu() - ora[k]).abs().max()) < TOL, k
        assert torch.equal(out[k], out_raw[k]), k
