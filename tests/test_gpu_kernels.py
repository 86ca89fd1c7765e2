"""GPU (-m gpu): every C-ABI kernel family against torch fp32 on the same seeded inputs, through the C ABI.
Tolerances: 3-pass bf16-split contractions 5e-5 relative-to-max (fp32-equivalent), 1xTF32 5e-3, elementwise 1e-5;
integer / index outputs bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import a2x_import

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return a2x_import.pkg("ops")


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


def rnd(g, *s):
    return torch.randn(*s, generator=g).cuda()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


CONV_CASES = [(1, 8, 16, 32, 32, 1, 1), (2, 12, 40, 64, 64, 3, 1), (2, 20, 44, 64, 128, 3, 2), (1, 25, 88, 128, 256, 3, 1),
              (2, 9, 21, 64, 128, 3, 2), (1, 10, 36, 384, 256, 1, 1), (1, 10, 36, 256, 32, 1, 1), (3, 7, 5, 32, 64, 3, 2)]


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(ops, case, split):
    n, h, w, cin, cout, k, s = case
    g = _g(1)
    x, wt = rnd(g, n, cin, h, w), rnd(g, cout, cin, k, k) * 0.1
    yref = F.conv2d(x, wt, stride=s, padding=k // 2)
    dy = rnd(g, *yref.shape)
    dxref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=s, padding=k // 2)
    dwref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=s, padding=k // 2)
    if split and (cin % 64 or cout % 64):
        pytest.skip("split operands need channel counts that are multiples of 64")
    pw = ops.pack_conv_weight(wt)
    if split:
        xs, dys, tol = ops.split(nhwc(x)), ops.split(nhwc(dy)), 5e-5
    else:
        xs, dys, tol = ops.Act(nhwc(x)), ops.Act(nhwc(dy)), 5e-3
    y = ops.Act(torch.empty_like(nhwc(yref)))
    ops.conv_fwd(xs, pw, k, s, y)
    dx = torch.empty_like(nhwc(x))
    ops.conv_dgrad(dys, pw, k, s, dx)
    dwp = torch.zeros(k * k, cout, cin, device="cuda")
    ops.conv_wgrad(xs, dys, k, s, dwp)
    dw = ops.unpack_conv_wgrad(dwp, cout, cin, k)
    assert rel(y.hi, nhwc(yref)) < tol
    assert rel(dx, nhwc(dxref)) < tol
    assert rel(dw, dwref) < tol


@pytest.mark.parametrize("case", [(1, 10, 36, 64, 128, 1), (1, 10, 18, 128, 128, 2), (2, 5, 9, 256, 128, 4)])
def test_deconv(ops, case):
    n, h, w, cin, cout, s = case
    g = _g(2)
    x = rnd(g, n, cin, h, w).requires_grad_(True)
    wt = (rnd(g, cin, cout, s, s) * 0.1).requires_grad_(True)
    yref = F.conv_transpose2d(x, wt, stride=s)
    dy = rnd(g, *yref.shape)
    yref.backward(dy)
    pw = ops.pack_deconv_weight(wt.detach())
    xs, dys = ops.split(nhwc(x.detach())), ops.split(nhwc(dy))
    # forward into a channel slice of a wider (concat) buffer, with fused affine + ReLU epilogue
    buf = torch.zeros(n, h * s, w * s, 384, device="cuda")
    scale = torch.rand(cout, generator=g).cuda() + 0.5
    shift = torch.randn(cout, generator=g).cuda() * 0.1
    ops.deconv_fwd(xs, pw, cout, s, ops.Act(buf[..., 128:128 + cout]), scale=scale, shift=shift, relu=True)
    want = F.relu(yref.detach() * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    assert rel(buf[..., 128:128 + cout], nhwc(want)) < 5e-5
    assert float(buf[..., :128].abs().max()) == 0.0 and float(buf[..., 128 + cout:].abs().max()) == 0.0
    dx = torch.empty_like(nhwc(x.detach()))
    ops.deconv_dgrad(dys, pw, s, dx)
    dwp = torch.zeros(s * s, cin, cout, device="cuda")
    ops.deconv_wgrad(xs, dys, s, dwp)
    assert rel(dx, nhwc(x.grad)) < 5e-5
    assert rel(ops.unpack_deconv_wgrad(dwp, cin, cout, s), wt.grad) < 5e-5


def test_split_representation(ops):
    x = rnd(_g(3), 1, 4, 8, 64) * 100
    a = ops.split(x)
    assert torch.equal(a.hi, x)                                            # the fp32 plane keeps the full value
    h = x.to(torch.bfloat16)
    assert torch.equal(a.b16[0], h)                                        # plane 0 = bf16(v)
    assert torch.equal(a.b16[1], (x - h.float()).to(torch.bfloat16))       # plane 1 = bf16(v - h16)
    back = a.b16[0].float() + a.b16[1].float()
    assert float((back - x).abs().max() / x.abs().max()) < 2.0 ** -16      # h16 + l16 ~ v to 2^-17


@pytest.mark.parametrize("case", [(4, 64, 9, 13), (3, 128, 5, 7), (5, 256, 4, 6), (1, 64, 3, 5), (16, 64, 2, 3)])
def test_attention_fusion(ops, case):
    n, C, H, W = case
    g = _g(4)
    x = rnd(g, n, H, W, C).requires_grad_(True)
    q = x.view(n, H * W, C).permute(1, 0, 2)
    ctx = torch.bmm(F.softmax(torch.bmm(q, q.transpose(1, 2)) / np.sqrt(C), -1), q)
    ref = ctx[:, 0].view(1, H, W, C)
    dout = rnd(g, 1, H, W, C)
    ref.backward(dout)
    out = ops.Act(torch.empty(1, H, W, C, device="cuda"))
    ops.att_fuse_fwd(x.detach(), out)
    dx = torch.empty_like(x)
    ops.att_fuse_bwd(x.detach(), dout, dx)
    assert rel(out.hi, ref.detach()) < 1e-5
    assert rel(dx, x.grad) < 1e-5


@pytest.mark.parametrize("case", [(3, 10, 12, 64), (2, 7, 9, 128), (1, 5, 6, 384)])
def test_batchnorm_relu_train(ops, case):
    N, H, W, C = case
    g = _g(5)
    z = rnd(g, N, H, W, C).requires_grad_(True)
    gam = (torch.rand(C, generator=g) + 0.5).cuda().requires_grad_(True)
    bet = (torch.randn(C, generator=g) * 0.1).cuda().requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    rm_ref, rv_ref = rm.clone(), rv.clone()
    for _ in range(3):  # the reference updates block-0 running stats three times per step
        y = F.relu(F.batch_norm(z.permute(0, 3, 1, 2), rm_ref, rv_ref, gam, bet, True, 0.01, 1e-3)).permute(0, 2, 3, 1)
    dy = rnd(g, N, H, W, C)
    y.backward(dy)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    scale, shift, mean, invstd = [torch.empty(C, device="cuda") for _ in range(4)]
    ops.channel_stats(z.detach(), sums)
    ops.bn_finalize(sums, N * H * W, gam.detach(), bet.detach(), 3, rm, rv, scale, shift, mean, invstd)
    yo = ops.Act.empty((N, H, W, C), "cuda", True)
    ops.affine_act(z.detach(), scale, shift, True, yo)
    bs = torch.zeros(ops.bn_bwd_sums_len(C), dtype=torch.float64, device="cuda")
    dz = ops.Act(torch.empty(N, H, W, C, device="cuda"))
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_relu_bwd(dy, z.detach(), scale, shift, mean, invstd, bs, dz, dg, db)
    assert rel(yo.value(), y.detach()) < 1e-5
    assert rel(dz.hi, z.grad) < 1e-5 and rel(dg, gam.grad) < 1e-5 and rel(db, bet.grad) < 1e-5
    assert rel(rm, rm_ref) < 1e-5 and rel(rv, rv_ref) < 1e-5


def test_communication_mask(ops):
    from oracle import w2c_oracle as O
    import random

    g = _g(6)
    N, H, W = 5, 20, 44
    record_len = [3, 2]
    psm = rnd(g, N, 14, H, W) * 2 - 4
    gw, gb = O.gaussian_filter_params(5, 1.0)
    sd = {"fusion_net.naive_communication.gaussian_filter.weight": gw, "fusion_net.naive_communication.gaussian_filter.bias": gb}
    comm = {"threshold": 0.01, "gaussian_smooth": {"k_size": 5, "c_sigma": 1.0}}
    heads = torch.zeros(N, H, W, 32, device="cuda")
    heads[..., :14] = nhwc(psm)
    conf, smooth, mask = [torch.empty(N, H, W, device="cuda") for _ in range(3)]
    ones = torch.empty(2, device="cuda")
    ss = torch.tensor([0, 3], dtype=torch.int32, device="cuda")
    sl = torch.tensor(record_len, dtype=torch.int32, device="cuda")
    for training in (False, True):
        random.seed(9)
        m_ref, rate_ref = O.communication(sd, comm, psm.cpu(), torch.tensor(record_len), training)
        ops.comm_confidence(heads, 14, conf)
        if training:
            random.seed(9)
            ks = [int(H * W * random.uniform(0, 1)) for _ in record_len]
            k_dev = torch.tensor([ks[0]] * 3 + [ks[1]] * 2, dtype=torch.int32, device="cuda")
            ops.comm_smooth_mask(conf, gw.cuda(), gb.cuda(), 5, N, H, W, 0.01, False, smooth, mask)
            ops.comm_topk_mask(smooth, N, H * W, k_dev, mask)
        else:
            ops.comm_smooth_mask(conf, gw.cuda(), gb.cuda(), 5, N, H, W, 0.01, True, smooth, mask)
        ops.comm_rate_ego(mask, H * W, 2, ss, sl, ones)
        assert torch.equal(mask.cpu().unsqueeze(1), m_ref), training                 # bit-exact mask
        rate = (ones.cpu() / (torch.tensor(record_len, dtype=torch.float32) * H * W)).sum() / 2
        assert abs(float(rate) - float(rate_ref)) < 1e-6


def test_topk_ties_and_extremes(ops):
    N, HW = 3, 1000
    smooth = torch.zeros(N, HW, device="cuda")
    smooth[1] = torch.arange(HW, device="cuda").float()
    smooth[2, ::2] = 1.0                                                          # 500 ties at the top
    k = torch.tensor([10, 0, 300], dtype=torch.int32, device="cuda")
    mask = torch.empty(N, HW, device="cuda")
    ops.comm_topk_mask(smooth, N, HW, k, mask)
    assert mask.sum(1).tolist() == [10.0, 0.0, 300.0]                              # exactly K ones, like torch.topk
    assert mask[0, :10].sum() == 10 and mask[2, 1::2].sum() == 0
    k = torch.tensor([HW, HW, HW], dtype=torch.int32, device="cuda")
    ops.comm_topk_mask(smooth, N, HW, k, mask)
    assert float(mask.min()) == 1.0


def test_det_loss_matches_oracle(ops):
    from oracle import w2c_oracle as O

    g = _g(7)
    B, H, W, A, K = 2, 12, 20, 2, 7
    heads = (rnd(g, B, H, W, 32)).requires_grad_(True)
    labels = O.make_labels(11, B, H, W, A, n_pos=25)
    labels["targets"].view(-1)[5] = float("nan")                                   # NaN targets are ignored (:61)
    nchw = heads.permute(0, 3, 1, 2)
    out = {"psm": nchw[:, :14].cpu(), "rm": nchw[:, 14:28].cpu(), "obj": nchw[:, 28:30].cpu()}
    tot, lr, lc, lo = O.point_pillar_loss_multiclass(out, labels, K, 1.0, 2.0)
    tot.backward()
    lab = {"targets": labels["targets"].float().cuda().contiguous(), "pos_equal_one": labels["pos_equal_one"].float().cuda().contiguous(),
           "class_ids": labels["class_ids"].int().cuda().contiguous()}
    loss3 = torch.zeros(3, dtype=torch.float64, device="cuda")
    npos = torch.zeros(B, device="cuda")
    dh = torch.empty(B, H, W, 32, device="cuda")
    ops.det_loss(heads.detach(), A, K, lab["targets"], lab["pos_equal_one"], lab["class_ids"], 1.0, 2.0, npos, dh, loss3)
    got = loss3.cpu().tolist()
    for a, b in zip(got, [float(lr), float(lc), float(lo)]):
        assert abs(a - b) < 1e-5 * max(1.0, abs(b))
    assert rel(dh, heads.grad) < 1e-5


def test_pfn_forward_backward_matches_oracle(ops):
    from oracle import voxelize as V, w2c_oracle as O

    g = _g(8)
    r = [-25.6, -12.8, -3, 25.6, 12.8, 1]
    vs = [0.4, 0.4, 4]
    per = [V.voxelize(V.mask_points(O.synth_points(50 + i, 4000, r, (10.0, 5.0)), r, ego_box=(i == 0)), r, vs) for i in range(2)]
    col = V.collate(per)
    vf, vn, vc = [torch.from_numpy(col[k]) for k in ("voxel_features", "voxel_num_points", "voxel_coords")]
    M = vf.shape[0]
    sd = {"p.pfn_layers.0.linear.weight": (torch.randn(64, 10, generator=g) * 0.3).requires_grad_(True),
          "p.pfn_layers.0.norm.weight": (torch.rand(64, generator=g) + 0.5).requires_grad_(True),
          "p.pfn_layers.0.norm.bias": (torch.randn(64, generator=g) * 0.1).requires_grad_(True),
          "p.pfn_layers.0.norm.running_mean": torch.zeros(64), "p.pfn_layers.0.norm.running_var": torch.ones(64)}
    for training in (False, True):
        buffers = {}
        pf, _ = O.pillar_vfe(sd, "p", vf, vn, vc, vs, r, training, buffers)
        canvas_ref = O.scatter(pf, vc, 128, 64, 2)
        geom = ops.pfn_geom(vs, r, 128, 64)
        w = sd["p.pfn_layers.0.linear.weight"].detach().cuda()
        gam, bet = sd["p.pfn_layers.0.norm.weight"].detach().cuda(), sd["p.pfn_layers.0.norm.bias"].detach().cuda()
        rm, rv = torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
        scale, shift, mean, invstd = [torch.empty(64, device="cuda") for _ in range(4)]
        vfc, vnc, vcc = vf.cuda(), vn.cuda(), vc.cuda()
        canvas = ops.Act(torch.zeros(2, 64, 128, 64, device="cuda"))
        amap = torch.tensor([0, 1], dtype=torch.int32, device="cuda")
        pout = torch.empty(M, 64, device="cuda")
        if training:
            mom = torch.empty(65, dtype=torch.float64, device="cuda")
            ops.pfn_moments(vfc, vnc, vcc, geom, mom)
            ops.pfn_stats_finalize(mom, M * 32, w, gam, bet, 1, rm, rv, scale, shift, mean, invstd)
            amax = torch.empty(M, 64, dtype=torch.uint8, device="cuda")
            ops.pfn_scatter(vfc, vnc, vcc, geom, w, scale, shift, amap, canvas, pillar_out=pout, amax=amax)
            assert rel(rm.cpu(), buffers["p.pfn_layers.0.norm.running_mean"]) < 1e-5
            assert rel(rv.cpu(), buffers["p.pfn_layers.0.norm.running_var"]) < 1e-5
        else:
            ops.bn_eval_affine(gam, bet, rm, rv, scale, shift)
            ops.pfn_scatter(vfc, vnc, vcc, geom, w, scale, shift, amap, canvas, pillar_out=pout)
        assert rel(pout.cpu(), pf.detach()) < 1e-5
        assert rel(canvas.hi.permute(0, 3, 1, 2).cpu(), canvas_ref.detach()) < 1e-5
        if training:
            dc = torch.randn(2, 64, 64, 128, generator=g)
            (canvas_ref * dc).sum().backward()
            acc = torch.empty(64 * 12, dtype=torch.float64, device="cuda")
            dw, dg, db = torch.empty(64, 10, device="cuda"), torch.empty(64, device="cuda"), torch.empty(64, device="cuda")
            ops.pfn_bwd(vfc, vnc, vcc, geom, w, scale, shift, mean, invstd, amap, dc.permute(0, 2, 3, 1).contiguous().cuda(),
                        amax, mom, M * 32, acc, dw, dg, db)
            assert rel(dw.cpu(), sd["p.pfn_layers.0.linear.weight"].grad) < 1e-4
            assert rel(dg.cpu(), sd["p.pfn_layers.0.norm.weight"].grad) < 1e-4
            assert rel(db.cpu(), sd["p.pfn_layers.0.norm.bias"].grad) < 1e-4


@pytest.mark.parametrize("case", [(2, 12, 40, 64, 64, 3, 1), (2, 21, 45, 64, 128, 3, 2), (1, 25, 88, 128, 256, 3, 1)])
def test_conv_epilogue_batch_statistics(ops, case):
    """BN batch statistics fused into the GEMM epilogue == a separate reduction over the written output
    (partial tiles and out-of-image rows excluded)."""
    n, h, w, cin, cout, k, s = case
    g = _g(12)
    x, wt = rnd(g, n, cin, h, w), rnd(g, cout, cin, k, k) * 0.1
    pw = ops.pack_conv_weight(wt)
    ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
    y = ops.Act(torch.empty(n, ho, wo, cout, device="cuda"))
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    ops.conv_fwd(ops.split(nhwc(x)), pw, k, s, y, stats=stats)
    ref = torch.cat([y.hi.double().sum((0, 1, 2)), (y.hi.double() ** 2).sum((0, 1, 2))])
    assert float((stats - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("case", [(2, 48, 80, 64, 64), (1, 21, 45, 128, 32), (3, 8, 8, 64, 96)])
def test_conv7x7_stride2_stem(ops, case):
    """7x7 stride-2 padding-3 conv (BevEncode.conv1, sub_modules/lss_submodule.py:318) as 49 taps over the four stride-2
    parity views, six accumulating launches; odd sizes, borders (taps reaching 3 pixels outside) == F.conv2d"""
    n, h, w, cin, cout = case
    g = _g(41)
    x, wt = rnd(g, n, cin, h, w), rnd(g, cout, cin, 7, 7) * 0.05
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = ops.Act(torch.full((n, ho, wo, cout), 7.0, device="cuda"))
    ops.conv_fwd(ops.split(nhwc(x)), ops.pack_conv_weight(wt), 7, 2, y)
    ref = F.conv2d(x.double(), wt.double(), stride=2, padding=3).permute(0, 2, 3, 1).float()
    assert ref.shape == y.hi.shape and rel(y.hi, ref) < 2e-5


def test_deconv_epilogue_batch_statistics(ops):
    g = _g(13)
    n, h, w, cin, cout, s = 2, 5, 9, 256, 128, 4
    x, wt = rnd(g, n, cin, h, w), rnd(g, cin, cout, s, s) * 0.1
    pw = ops.pack_deconv_weight(wt)
    y = ops.Act(torch.empty(n, h * s, w * s, cout, device="cuda"))
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    ops.deconv_fwd(ops.split(nhwc(x)), pw, cout, s, y, stats=stats)
    ref = torch.cat([y.hi.double().sum((0, 1, 2)), (y.hi.double() ** 2).sum((0, 1, 2))])
    assert float((stats - ref).abs().max() / ref.abs().max()) < 1e-5


def test_mask_compact_decompact_roundtrip(ops):
    """sparse feature select (warp-ballot compaction) followed by the pointer-table scatter == x * mask, per agent"""
    g = _g(11)
    n, H, W, C = 3, 13, 21, 64
    hw = H * W
    x = rnd(g, n, H, W, C)
    mask = (torch.rand(n, H, W, generator=g) > 0.7).float().cuda()
    total = 64 + (hw + 3) // 4 * 4 + hw * C
    bufs = torch.zeros(n, total, device="cuda")
    for a in range(n):
        hdr = bufs[a, :64].view(torch.int32)
        idx = bufs[a, 64:64 + (hw + 3) // 4 * 4].view(torch.int32)
        vals = bufs[a, 64 + (hw + 3) // 4 * 4:]
        ops.mask_compact(x[a:a + 1], mask[a], a == 0, hdr, idx, vals)
        assert int(hdr[1]) == int(mask[a].sum())                          # communication-rate numerator
        assert int(hdr[0]) == (hw if a == 0 else int(mask[a].sum()))      # the ego sends every cell
    table = torch.tensor([bufs[a].data_ptr() for a in range(n)], dtype=torch.int64, device="cuda")
    dst = torch.full((n, H, W, C), 5.0, device="cuda")
    ops.mask_decompact_ptrs(table, 64 * 4, (64 + (hw + 3) // 4 * 4) * 4, n, dst)
    want = x * mask[..., None]
    want[0] = x[0]
    assert torch.equal(dst, want)


def test_dgrad_fused_bn_bwd_reduction(ops):
    """the BN+ReLU backward reduction accumulated in the data-gradient epilogue == the separate reduce pass over (dx, z)"""
    g = _g(12)
    n, h, w, cin, cout = 2, 19, 37, 128, 256
    dy = rnd(g, n, cout, h, w)
    wt = rnd(g, cout, cin, 3, 3) * 0.1
    z = rnd(g, n, h, w, cin)
    scale, shift = torch.rand(cin, generator=g).cuda() + 0.5, torch.randn(cin, generator=g).cuda() * 0.3
    mean, invstd = torch.randn(cin, generator=g).cuda() * 0.2, torch.rand(cin, generator=g).cuda() + 0.5
    pw = ops.pack_conv_weight(wt)
    dys = ops.split(nhwc(dy))
    dx_ref, dx = torch.empty(n, h, w, cin, device="cuda"), torch.empty(n, h, w, cin, device="cuda")
    ops.conv_dgrad(dys, pw, 3, 1, dx_ref)
    sums = torch.zeros(ops.bn_bwd_sums_len(cin), dtype=torch.float64, device="cuda")
    ops.conv_dgrad(dys, pw, 3, 1, dx, bn_stats=(z, scale, shift, mean, invstd, sums))
    sums = sums[:2 * cin]
    assert torch.equal(dx, dx_ref)
    gate = (z * scale + shift > 0).double()
    gg = dx_ref.double() * gate
    want = torch.cat([gg.sum((0, 1, 2)), (gg * ((z - mean) * invstd).double()).sum((0, 1, 2))])
    assert float((sums - want).abs().max() / want.abs().max()) < 1e-5
