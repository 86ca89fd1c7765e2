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


def test_hgt_backward_through_fold(ops):
    """d(typed q/k/v linears, relation tensors, input) of HGTCavAttention (identity output projection) == torch autograd
    on the oracle: hgt_attention_bwd -> fp32 dgrad/wgrad of the folded projection (torch) -> hgt_fold_bwd"""
    g = torch.Generator().manual_seed(7)
    n, H, W, C, heads, dh = 4, 4, 5, 256, 8, 32
    pre = "f"
    sd = {}
    for name in ("q_linears", "k_linears", "v_linears"):
        for t in range(2):
            sd["%s.%s.%d.weight" % (pre, name, t)] = (torch.randn(C, C, generator=g, dtype=torch.float64) * 0.05).requires_grad_(True)
            sd["%s.%s.%d.bias" % (pre, name, t)] = (torch.randn(C, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True)
    for t in range(2):
        sd["%s.a_linears.%d.weight" % (pre, t)] = torch.eye(C, dtype=torch.float64)
        sd["%s.a_linears.%d.bias" % (pre, t)] = torch.zeros(C, dtype=torch.float64)
    sd[pre + ".relation_att"] = (torch.randn(4, heads, dh, dh, generator=g, dtype=torch.float64) * 0.2).requires_grad_(True)
    sd[pre + ".relation_msg"] = (torch.randn(4, heads, dh, dh, generator=g, dtype=torch.float64) * 0.2).requires_grad_(True)
    x = torch.randn(1, n, H, W, C, generator=g, dtype=torch.float64).requires_grad_(True)
    types = [0, 1, 1, 0]
    prior = torch.zeros(1, n, H, W, 3, dtype=torch.float64)
    prior[0, :, :, :, 2] = torch.tensor(types, dtype=torch.float64)[:, None, None]
    mask = (torch.rand(1, H, W, 1, n, generator=g) > 0.3).double()
    mask[..., 0] = 1.0
    dout = torch.randn(n, H, W, C, generator=g, dtype=torch.float64)
    VO.hgt_attention(sd, pre, x, mask, prior, heads, dh)[0].backward(dout)
    # CUDA path
    f32 = lambda t: t.detach().float().cuda()
    pair = lambda nme, s: (f32(sd["%s.%s.0.%s" % (pre, nme, s)]), f32(sd["%s.%s.1.%s" % (pre, nme, s)]))
    wf, bf = torch.empty(2, 5 * C, C, device="cuda"), torch.empty(2, 5 * C, device="cuda")
    ops.hgt_fold(pair("q_linears", "weight"), pair("q_linears", "bias"), pair("k_linears", "weight"), pair("k_linears", "bias"),
                 pair("v_linears", "weight"), pair("v_linears", "bias"), f32(sd[pre + ".relation_att"]), f32(sd[pre + ".relation_msg"]),
                 heads, wf, bf)
    xc = f32(x)[0]
    qkv = torch.stack([F.linear(xc[a], wf[types[a]], bf[types[a]]) for a in range(n)])
    km = mask[0, :, :, 0, :].permute(2, 0, 1).contiguous().float().cuda()
    tdev = torch.tensor(types, dtype=torch.int32).cuda()
    dqkv = torch.empty_like(qkv)
    ops.hgt_attention_bwd(qkv, tdev, km, dout.float().cuda(), heads, dh, dqkv)
    dwf, dbf, dx = torch.zeros_like(wf), torch.zeros_like(bf), torch.zeros(n, H, W, C, device="cuda")
    for a in range(n):                                     # the projection's own backward: plain fp32 torch here
        g2 = dqkv[a].reshape(-1, 5 * C)
        dwf[types[a]] += g2.t() @ xc[a].reshape(-1, C)
        dbf[types[a]] += g2.sum(0)
        dx[a] = (g2 @ wf[types[a]]).reshape(H, W, C)
    outs = {k: [torch.empty(C, C, device="cuda") for _ in range(2)] for k in ("q", "k", "v")}
    outb = {k: [torch.empty(C, device="cuda") for _ in range(2)] for k in ("q", "k", "v")}
    dA, dM = torch.empty(4, heads, dh, dh, device="cuda"), torch.empty(4, heads, dh, dh, device="cuda")
    ops.hgt_fold_bwd(dwf, dbf, pair("k_linears", "weight"), pair("k_linears", "bias"), pair("v_linears", "weight"),
                     pair("v_linears", "bias"), f32(sd[pre + ".relation_att"]), f32(sd[pre + ".relation_msg"]), heads,
                     outs["q"], outb["q"], outs["k"], outb["k"], outs["v"], outb["v"], dA, dM)
    assert rel(dx.cpu().double(), x.grad[0]) < 5e-5
    assert rel(dA.cpu().double(), sd[pre + ".relation_att"].grad) < 5e-5
    assert rel(dM.cpu().double(), sd[pre + ".relation_msg"].grad) < 5e-5
    for k, nme in (("q", "q_linears"), ("k", "k_linears"), ("v", "v_linears")):
        for t in range(2):
            assert rel(outs[k][t].cpu().double(), sd["%s.%s.%d.weight" % (pre, nme, t)].grad) < 5e-5, (k, t)
            assert rel(outb[k][t].cpu().double(), sd["%s.%s.%d.bias" % (pre, nme, t)].grad) < 5e-5, (k, t)


def test_split_attn_and_rte_backward(ops):
    g = torch.Generator().manual_seed(8)
    n, H, W, C = 3, 6, 8, 256
    wins = [torch.randn(1, n, H, W, C, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(3)]
    sd = {"s.fc1.weight": (torch.randn(C, C, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True),
          "s.bn1.weight": (torch.rand(C, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True),
          "s.bn1.bias": (torch.randn(C, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True),
          "s.fc2.weight": (torch.randn(3 * C, C, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True)}
    dx = torch.randn(n, H, W, C, generator=g, dtype=torch.float64)
    VO.split_attn(sd, "s", wins)[0].backward(dx)
    f32 = lambda t: t.detach().float().cuda().contiguous()
    w = [f32(t[0]) for t in wins]
    x = torch.zeros(n, H, W, C, device="cuda")
    sums, wts = torch.empty(n, C, device="cuda"), torch.empty(n, 3, C, device="cuda")
    ops.split_attn_fuse(w[0], w[1], w[2], f32(sd["s.fc1.weight"]), f32(sd["s.bn1.weight"]), f32(sd["s.bn1.bias"]),
                        f32(sd["s.fc2.weight"]), sums, wts, x)              # forward leaves the pooled sums / weights
    d = [ops.Act.empty((n, H, W, C), "cuda", True) for _ in range(3)]
    dfc1, dg, db, dfc2 = (torch.zeros(C, C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"),
                          torch.zeros(3 * C, C, device="cuda"))
    ops.split_attn_bwd(dx.float().cuda(), w[0], w[1], w[2], f32(sd["s.fc1.weight"]), f32(sd["s.bn1.weight"]), f32(sd["s.bn1.bias"]),
                       f32(sd["s.fc2.weight"]), sums, wts, torch.empty(n, 3, C, device="cuda"), torch.empty(n, C, device="cuda"),
                       d[0], d[1], d[2], dfc1, dg, db, dfc2)
    for r in range(3):
        assert rel(d[r].hi.cpu().double(), wins[r].grad[0]) < 5e-5, r
    assert rel(dfc1.cpu().double(), sd["s.fc1.weight"].grad) < 1e-4 and rel(dfc2.cpu().double(), sd["s.fc2.weight"].grad) < 1e-4
    assert rel(dg.cpu().double(), sd["s.bn1.weight"].grad) < 1e-4 and rel(db.cpu().double(), sd["s.bn1.bias"].grad) < 1e-4
    # RTE
    emb = torch.randn(100, C, generator=g, dtype=torch.float64).requires_grad_(True)
    lw = (torch.randn(C, C, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True)
    lb = torch.randn(C, generator=g, dtype=torch.float64).requires_grad_(True)
    idx = torch.tensor([0, 4, 4])
    xx = torch.randn(n, H, W, C, generator=g, dtype=torch.float64)
    (xx + F.linear(emb[idx], lw, lb)[:, None, None, :]).backward(dx)
    sums2 = torch.zeros(n, 2 * C, dtype=torch.float64, device="cuda")
    dxc = dx.float().cuda()
    for a in range(n):
        ops.channel_stats(dxc[a:a + 1], sums2[a])
    dW, dB, dE = torch.zeros(C, C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(100, C, device="cuda")
    ops.rte_bwd(sums2, emb.detach().float().cuda(), idx.int().cuda(), lw.detach().float().cuda(), dW, dB, dE)
    assert rel(dW.cpu().double(), lw.grad) < 5e-5 and rel(dB.cpu().double(), lb.grad) < 5e-5 and rel(dE.cpu().double(), emb.grad) < 5e-5


def test_train_step_matches_oracle_autograd():
    """V2X-ViT training step (train-mode BatchNorm, nn.Dropout off; the dropout-on step is the next test): loss and every parameter gradient against torch autograd
    through the oracle (pinned to the real reference in eval mode). Fusion-network / head gradients are tight; encoder
    gradients pass ReLU / max gates (see tests/test_gpu_model.py) -> norm-wise bounds. `prior_feed` is unused: zero grads."""
    import json

    import a2x_import
    import w2c_common as C
    from oracle import w2c_oracle as O

    M = a2x_import.pkg("opencood.models.airv2x_v2xvit")
    cfg, gold = VC.load_small()
    args = json.loads(json.dumps(cfg["model_args"]))
    model = M.Airv2xV2XVit(args)
    sd = VC.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().train()
    dd = VC.golden_scene(cfg, gold)
    H, W = gold["eval_psm"].shape[2:]
    labels = O.make_labels(5, 1, H, W, args["anchor_number"])
    loss3 = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"],
                             dropout="off")
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
         for k, v in sd.items()}
    torch.set_num_threads(8)
    out, _ = VO.v2xvit_forward(p, args, dd, training=True)
    loss = O.point_pillar_loss_multiclass(out, labels, args["num_class"], cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])[0]
    loss.backward()
    assert abs(float(loss3.sum()) - float(loss.detach())) < 1e-3 * abs(float(loss.detach()))
    errs = {}
    for n, q in model.named_parameters():
        ref = p[n].grad
        if ref is None:
            assert float(q.grad.abs().max()) == 0.0, n
            continue
        errs[n] = float((q.grad.cpu() - ref).norm() / (ref.norm() + 1e-30))
    print(sorted(errs.items(), key=lambda kv: -kv[1])[:12])
    fusion = {n: e for n, e in errs.items() if n.startswith("fusion_net") or "head" in n}
    assert len(fusion) > 100 and max(fusion.values()) < 2e-2, sorted(fusion.items(), key=lambda kv: -kv[1])[:5]
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
    # a second step reuses every buffer (stale-state check): same loss, same gradients
    g1 = {n: q.grad.clone() for n, q in model.named_parameters()}
    loss3b = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"],
                              dropout="off")
    assert torch.allclose(loss3b, loss3, rtol=1e-6)
    worst = max(float((q.grad - g1[n]).norm() / (g1[n].norm() + 1e-30)) for n, q in model.named_parameters())
    assert worst < 1e-3, worst


def test_ragged_multi_scene_batch_matches_oracle(small):
    """B = 3 ragged scenes (4, 2 and 3 agents, mixed agent types, per-scene poses / delays): valid-agent regrouping,
    per-scene HGT attention with typed projections, STTF per agent — against the oracle on the CPU"""
    import math

    import w2c_common as C

    cfg, gold, model = small
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    scenes = [["vehicle", "vehicle", "rsu", "drone"], ["vehicle", "drone"], ["vehicle", "rsu", "rsu"]]
    pre = dict(cfg["preprocess"])
    pre["args"] = dict(pre["args"])
    pre["args"]["max_voxel_test"] = pre["args"]["max_voxel_train"]
    dd, _ = C.make_batch(pre, scenes, 4000, 43, pre["args"]["max_voxel_train"])
    L = sum(cfg["model_args"]["max_cav"].values())
    g = np.random.default_rng(3)
    prior = torch.zeros(len(scenes), L, 3)
    scm = torch.eye(4, dtype=torch.float64).repeat(len(scenes), L, 1, 1)
    for b, agents in enumerate(scenes):
        for i, t in enumerate(agents):
            prior[b, i] = torch.tensor([0.1 * i, float((i + b) % 3), 1.0 if t == "rsu" else 0.0])
            if i:
                a = g.uniform(-0.3, 0.3)
                scm[b, i, :2, :2] = torch.tensor([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]], dtype=torch.float64)
                scm[b, i, 0, 3], scm[b, i, 1, 3] = g.uniform(-6, 6), g.uniform(-3, 3)
    dd["prior_encoding"], dd["spatial_correction_matrix"] = prior, scm
    torch.set_num_threads(8)
    with torch.no_grad():
        ora, _ = VO.v2xvit_forward(sd, cfg["model_args"], dd, training=False)
        out = model(C.to_device(dd, "cuda"))
    assert out["psm"].shape[0] == 3
    for k in ("psm", "rm", "obj"):
        assert float((out[k].cpu() - ora[k]).abs().max()) < TOL, k
    assert out["comm_rate"] == ora["comm_rate"]


def test_train_step_with_dropout_matches_oracle_under_identical_masks():
    """The shipped yaml trains with nn.Dropout(0.3) in HGTCavAttention (hmsa.py:155), the window branches (mswin.py:47)
    and the feed forward (base_transformer.py:22,24). The kernels draw counter-based masks; exported (a2x_dropout_mask) and
    fed to the oracle, loss and gradients must agree like in the dropout-off test. Valid agents only (max_cav = #agents),
    so the mask layout [N, H, W, C] is the oracle's (1, L, H, W, C)."""
    import json

    import a2x_import
    import w2c_common as C
    from oracle import w2c_oracle as O

    M = a2x_import.pkg("opencood.models.airv2x_v2xvit")
    ops = a2x_import.pkg("ops")
    cfg, gold = VC.load_small()
    args = json.loads(json.dumps(cfg["model_args"]))
    model = M.Airv2xV2XVit(args)
    ps = tuple(model._dropouts())
    assert ps == (0.3, 0.3, 0.3)                       # the shipped yaml values
    sd = VC.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().train()
    dd = VC.golden_scene(cfg, gold)
    H, W = gold["eval_psm"].shape[2:]
    labels = O.make_labels(5, 1, H, W, args["anchor_number"])
    loss3 = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"], dropout=991).clone()
    loss_off = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"], dropout="off").clone()
    assert abs(float(loss3.sum()) - float(loss_off.sum())) > 1e-4 * abs(float(loss_off.sum()))
    model.load_state_dict(sd)
    loss3 = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"], dropout=991).clone()
    dc, dw, df = model.last_dropout
    n_agents = len([a for a in gold["agents"]])
    enc = args["transformer"]["encoder"]
    Cc, mlp, depth = enc["cav_att_config"]["dim"], enc["feed_forward"]["mlp_dim"], enc["depth"]
    n_tok = n_agents * H * W
    assert dc.n_sites == depth and dw.n_sites == 1000 + 3 * depth and df.n_sites == 2000 + 2 * depth
    masks = {"cav": [ops.dropout_mask(n_tok * Cc, dc, s).cpu() for s in range(depth)],
             "win": [ops.dropout_mask(n_tok * Cc, dw, 1000 + s).cpu() for s in range(3 * depth)],
             "ffn": [ops.dropout_mask(n_tok * (mlp if s % 2 == 0 else Cc), df, 2000 + s).cpu() for s in range(2 * depth)]}
    a5 = json.loads(json.dumps(args))
    agents = [str(a) for a in gold["agents"]]
    a5["max_cav"] = {t: sum(1 for a in agents if a == t) for t in O.AGENT_TYPES}
    dd5 = dict(dd)
    dd5["prior_encoding"], dd5["spatial_correction_matrix"] = dd["prior_encoding"][:, :n_agents], dd["spatial_correction_matrix"][:, :n_agents]
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    torch.set_num_threads(8)
    out, _ = VO.v2xvit_forward(p, a5, dd5, training=True, dropouts=VO.MaskedDropouts(ps, masks))
    loss = O.point_pillar_loss_multiclass(out, labels, args["num_class"], cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])[0]
    loss.backward()
    assert abs(float(loss3.sum()) - float(loss.detach())) < 1e-3 * abs(float(loss.detach()))
    errs = {}
    for n, q in model.named_parameters():
        if p[n].grad is not None:
            errs[n] = float((q.grad.cpu() - p[n].grad).norm() / (p[n].grad.norm() + 1e-30))
    fusion = {n: e for n, e in errs.items() if n.startswith("fusion_net") or "head" in n}
    assert len(fusion) > 100 and max(fusion.values()) < 2e-2, sorted(fusion.items(), key=lambda kv: -kv[1])[:5]
    assert float(np.median(list(fusion.values()))) < 2e-3
