"""CPU (-m "not gpu"): host logic of the criterion shims (`airv2x-perception_b200/det_loss.py`): recognising the models'
fused NHWC logit tensor behind the three NCHW-shaped outputs (read in place, no packing copy) and falling back to a
packed copy for anything else; no kernel is launched here."""
import torch


def test_fused_view_recognises_channel_slices_of_one_nhwc_tensor(pkg):
    import a2x_import

    D = a2x_import.pkg("det_loss")
    B, H, W, A, K, cs = 2, 5, 7, 2, 7, 64
    nc, nr = A * K, 7 * A
    heads = torch.randn(B, H, W, cs)
    nchw = heads.permute(0, 3, 1, 2)
    psm, rm, obj = nchw[:, :nc], nchw[:, nc:nc + nr], nchw[:, nc + nr:nc + nr + A]
    v = D._fused_view(psm, rm, obj)
    assert v is not None and v.shape == (B, H, W, nc + nr + A) and v.stride() == (H * W * cs, W * cs, cs, 1)
    assert v.data_ptr() == heads.data_ptr() and torch.equal(v, heads[..., :nc + nr + A])
    # legacy: psm | rm only
    v2 = D._fused_view(nchw[:, :A], nchw[:, A:A + nr], None)
    assert v2 is not None and v2.shape == (B, H, W, A + nr) and torch.equal(v2, heads[..., :A + nr])
    # a view that starts inside the tensor (a storage offset)
    sub = heads[1:]
    n2 = sub.permute(0, 3, 1, 2)
    v3 = D._fused_view(n2[:, :nc], n2[:, nc:nc + nr], n2[:, nc + nr:nc + nr + A])
    assert v3 is not None and torch.equal(v3, sub[..., :nc + nr + A])


def test_fused_view_rejects_everything_else(pkg):
    import a2x_import

    D = a2x_import.pkg("det_loss")
    B, H, W, A, K = 2, 5, 7, 2, 7
    nc, nr = A * K, 7 * A
    sep = (torch.randn(B, nc, H, W), torch.randn(B, nr, H, W), torch.randn(B, A, H, W))
    assert D._fused_view(*sep) is None                                   # three separate NCHW tensors
    heads = torch.randn(B, H, W, 64)
    nchw = heads.permute(0, 3, 1, 2)
    psm, rm, obj = nchw[:, :nc], nchw[:, nc:nc + nr], nchw[:, nc + nr:nc + nr + A]
    assert D._fused_view(psm, obj, rm) is None                           # wrong channel order
    assert D._fused_view(psm, rm, nchw[:, nc + nr + 1:nc + nr + 1 + A]) is None      # a gap between the slices
    assert D._fused_view(psm.double(), rm.double(), obj.double()) is None            # not fp32
    other = torch.randn(B, H, W, 64).permute(0, 3, 1, 2)
    assert D._fused_view(psm, other[:, nc:nc + nr], obj) is None         # slices of different tensors
    assert D._fused_view(psm.contiguous(), rm.contiguous(), obj.contiguous()) is None
    tight = torch.randn(B, H, W, nc + nr).permute(0, 3, 1, 2)            # pixel stride too small for the objectness slice
    assert D._fused_view(tight[:, :nc], tight[:, nc:], obj) is None


def test_backward_returns_one_gradient_per_forward_argument(pkg):
    """the Function's forward takes 10 arguments; its backward must hand back 10 entries in both layouts"""
    import inspect

    import a2x_import

    D = a2x_import.pkg("det_loss")
    n_args = len(inspect.signature(D.FusedDetLoss.forward).parameters) - 1          # minus ctx

    class Ctx:
        pass
    for legacy, split in ((False, [14, 14, 2]), (True, [2, 14])):
        ctx = Ctx()
        ctx.saved_tensors = (torch.randn(1, 3, 4, sum(split)),)
        ctx.split, ctx.legacy = split, legacy
        out = D.FusedDetLoss.backward(ctx, torch.tensor(2.0, dtype=torch.float64), None)
        assert len(out) == n_args == 10
        grads = [g for g in out[:3] if g is not None]
        assert [g.shape[1] for g in grads] == split and all(g.shape == (1, c, 3, 4) for g, c in zip(grads, split))
        assert torch.allclose(torch.cat(grads, 1).permute(0, 2, 3, 1), 2.0 * ctx.saved_tensors[0])
        assert all(g is None for g in out[3:])


def test_label_dict_of_boxes_is_assigned_lazily_and_label_maps_pass_through(pkg, monkeypatch):
    """`criterion(output_dict, batch["ego"]["label_dict"])` with this repo's dataset: the label_dict holds the padded boxes and
    the criterion assigns the anchor targets itself (one TargetAssigner per postprocess block and device, also after the
    reference's `to_device` rebuilt the dict); a label_dict with ready maps is used as is. The assigner is stubbed here."""
    import a2x_import

    D = a2x_import.pkg("det_loss")
    L = a2x_import.pkg("labels")
    made, calls = [], []

    class Stub:
        def __init__(self, params, device):
            made.append((params["max_num"], str(device)))

        def __call__(self, box, mask, cls):
            calls.append((box, mask, cls))
            return {"targets": "T", "pos_equal_one": "P", "class_ids": "C"}
    monkeypatch.setattr(L, "TargetAssigner", Stub)
    crit = D._Criterion({"cls_weight": 1.0, "reg": 2.0})
    ready = {"targets": 1, "pos_equal_one": 2, "class_ids": 3}
    assert crit._targets(ready, torch.device("cpu")) is ready and not made
    box, mask, cls = torch.zeros(1, 300, 7), torch.zeros(1, 300), torch.zeros(1, 300, dtype=torch.int64)
    lazy = {"object_bbx_center": box, "object_bbx_mask": mask, "object_class_ids": cls, "postprocess": {"max_num": 300, "order": "hwl"}}
    out = crit._targets(lazy, torch.device("cpu"))
    assert out == {"targets": "T", "pos_equal_one": "P", "class_ids": "C"} and calls[-1][0] is box and calls[-1][2] is cls
    rebuilt = {k: (dict(v) if isinstance(v, dict) else v) for k, v in lazy.items()}       # what to_device() hands over
    crit._targets(rebuilt, torch.device("cpu"))
    assert made == [(300, "cpu")] and len(calls) == 2                                     # same block: the assigner is reused
    crit._targets(dict(lazy, postprocess={"max_num": 100, "order": "hwl"}), torch.device("cpu"))
    assert made == [(300, "cpu"), (100, "cpu")]


def test_trainer_label_dispatch(pkg):
    """Trainer.labels: ready label maps pass through, a label_dict of boxes (this repo's dataset) goes to the GPU assigner"""
    import a2x_import

    TL = a2x_import.pkg("train_loop")

    class Self:
        def assigner(self, box, mask, cls):
            return ("assigned", box, mask, cls)
    ready = {"label_dict": {"targets": 1, "pos_equal_one": 2, "class_ids": 3}}
    assert TL.Trainer.labels(Self(), ready) is ready["label_dict"]
    boxes = {"object_bbx_center": "b", "object_bbx_mask": "m", "object_class_ids": "c"}
    assert TL.Trainer.labels(Self(), dict(boxes, label_dict=dict(boxes, postprocess={}))) == ("assigned", "b", "m", "c")
    assert TL.Trainer.labels(Self(), boxes) == ("assigned", "b", "m", "c")
