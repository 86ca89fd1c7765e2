"""The reference's criterion objects on the fused loss kernel (`train_utils.create_loss`, tools/train_utils.py:328-368,
finds `opencood.loss.<core_method>` by module path + class name; `tools/train.py:222-226` then calls
`criterion(output_dict, label_dict)`, `loss.backward()` and `criterion.logging(...)`).

`FusedDetLoss` evaluates value AND gradient w.r.t. the head logits in one launch (`a2x_det_loss` / `a2x_det_loss_legacy`,
csrc/loss.cu) and hands autograd the stored gradient, so the reference's own loop trains on the kernels without the ~20
elementwise launches and three `.item()` syncs of `PointPillarLossMultiClass.forward`
(loss/point_pillar_loss_multiclass.py:96-179). The models of this repo return `psm` / `rm` / `obj` as channel slices of ONE
NHWC logit tensor; that tensor is then read in place (no packing copy). No CPU path."""
import torch

from . import ops


def _fused_view(psm, rm, obj):
    """the [B,H,W,C] NHWC tensor the three NCHW-shaped outputs are channel slices of, or None"""
    parts = [t for t in (psm, rm, obj) if t is not None]
    B, _, H, W = psm.shape
    cs = psm.stride(3)
    want = (H * W * cs, 1, W * cs, cs)
    off, esz = 0, psm.element_size()
    for t in parts:
        if t.dtype != torch.float32 or tuple(t.stride()) != want or t.shape[0] != B or tuple(t.shape[2:]) != (H, W) \
                or t.data_ptr() != psm.data_ptr() + off * esz:
            return None
        off += t.shape[1]
    if off > cs:
        return None
    return torch.as_strided(psm, (B, H, W, off), (H * W * cs, W * cs, cs, 1), psm.storage_offset())


class FusedDetLoss(torch.autograd.Function):
    """(psm, rm, obj | None) NCHW logits + label tensors -> total loss (float64 scalar, like the reference's sum with its
    fp64 labels) and the (reg, cls, obj) terms; backward = the gradient the kernel wrote, scaled by the incoming one."""

    @staticmethod
    def forward(ctx, psm, rm, obj, targets, pos, class_ids, A, K, cls_weight, reg_coe):
        if not psm.is_cuda:
            raise RuntimeError("the fused detection loss (B200) needs CUDA tensors; there is no CPU path")
        legacy = obj is None
        parts = [psm, rm] if legacy else [psm, rm, obj]
        heads = _fused_view(psm, rm, obj)
        if heads is None:
            heads = torch.cat([t.float() for t in parts], dim=1).permute(0, 2, 3, 1).contiguous()
        B, H, W = heads.shape[:3]
        dev = heads.device
        dheads = torch.empty((B, H, W, heads.shape[3]), dtype=torch.float32, device=dev)
        loss3 = torch.empty(3, dtype=torch.float64, device=dev)
        npos = torch.empty(B, dtype=torch.float32, device=dev)
        tg = targets.to(device=dev, dtype=torch.float32).reshape(B, H, W, 7 * A).contiguous()
        ps = pos.to(device=dev, dtype=torch.float32).reshape(B, H, W, A).contiguous()
        ci = None if legacy else class_ids.to(device=dev, dtype=torch.int32).reshape(B, H, W, A).contiguous()
        ops.det_loss(heads, A, K, tg, ps, ci, float(cls_weight), float(reg_coe), npos, dheads, loss3, legacy=legacy)
        ctx.save_for_backward(dheads)
        ctx.split = [t.shape[1] for t in parts]
        ctx.legacy = legacy
        ctx.mark_non_differentiable(loss3)
        return loss3.sum(), loss3

    @staticmethod
    def backward(ctx, g_total, _g3):
        (dheads,) = ctx.saved_tensors
        d = (dheads * g_total.to(torch.float32)).permute(0, 3, 1, 2)
        outs, c0 = [], 0
        for c in ctx.split:
            outs.append(d[:, c0:c0 + c])
            c0 += c
        if ctx.legacy:
            outs.append(None)
        return (*outs, None, None, None, None, None, None, None)


class _Criterion(torch.nn.Module):
    """shared shell of the two criterion classes: the reference's constructor keys, `loss_dict`, `logging`"""
    legacy = False

    def __init__(self, args):
        super().__init__()
        self.cls_weight = args["cls_weight"]
        self.reg_coe = args["reg"]
        self.loss_dict = {}
        self._terms = {}
        self._assigner = None

    def _targets(self, target_dict, device):
        """`label_dict` of this repo's dataset = the padded boxes (+ the yaml's postprocess block) instead of label maps: the
        anchor targets are assigned here, on the GPU (`labels.TargetAssigner`, one per postprocess block and device), so
        the reference's loop — `criterion(output_dict, batch_data["ego"]["label_dict"])`, tools/train.py:222-226 — runs
        unmodified on this dataset. A `label_dict` that already holds `targets` (the reference's dataset) passes through."""
        if "targets" in target_dict:
            return target_dict
        from .labels import TargetAssigner
        import json
        key = (str(device), json.dumps(target_dict["postprocess"], sort_keys=True, default=str))   # content: to_device() rebuilds the dict
        if self._assigner is None or self._assigner[0] != key:
            self._assigner = (key, TargetAssigner(target_dict["postprocess"], device))
        return self._assigner[1](target_dict["object_bbx_center"], target_dict["object_bbx_mask"],
                                 target_dict["object_class_ids"])

    def _call(self, output_dict, target_dict, prefix, K):
        target_dict = self._targets(target_dict, output_dict["psm" + prefix].device)
        psm, rm = output_dict["psm" + prefix], output_dict["rm" + prefix]
        obj = None if self.legacy else output_dict["obj" + prefix]
        A = rm.shape[1] // 7
        total, loss3 = FusedDetLoss.apply(psm, rm, obj, target_dict["targets"], target_dict["pos_equal_one"],
                                          None if self.legacy else target_dict["class_ids"], A, K, self.cls_weight,
                                          self.reg_coe)
        # device tensors; read lazily by logging() so the step itself has no host sync (the reference calls .item() x3)
        self._terms["total_loss" + prefix] = total.detach()
        self._terms["reg_loss" + prefix] = loss3[0]
        self._terms["conf_loss" + prefix] = loss3[1]
        return total

    def logging(self, epoch, batch_id, batch_len, writer=None):
        """same message and tensorboard scalars as the reference's `logging` (point_pillar_loss_multiclass.py:300-333)"""
        self.loss_dict.update({k: float(v) for k, v in self._terms.items()})
        tot = [v for k, v in self.loss_dict.items() if "total_loss" in k]
        tot = sum(tot) if len(tot) > 1 else tot[0]
        msg = "[epoch {}][{}/{}], || Loss: {:.2f} ||".format(epoch, batch_id + 1, batch_len, tot)
        for k, v in self.loss_dict.items():
            msg += "{}: {:.2f} | ".format(k.replace("_loss", "").replace("_single", ""), v)
        if writer is not None:
            for k, v in self.loss_dict.items():
                writer.add_scalar(k, v, epoch * batch_len + batch_id)
        return msg
