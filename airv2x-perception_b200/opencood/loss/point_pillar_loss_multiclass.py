"""`opencood.loss.point_pillar_loss_multiclass` on the B200 path: `PointPillarLossMultiClass(args)` with the reference's
constructor keys (`cls_weight`, `reg`, `num_class`), `forward(output_dict, target_dict, prefix="")` and `logging`
(loss/point_pillar_loss_multiclass.py:77-179, :300-333): sigmoid focal (alpha .25, gamma 2) / #positives, smooth-L1
(beta 1/9) with the sin-difference on yaw, BCE on objectness — value and gradient from ONE kernel (csrc/loss.cu)."""
from ...det_loss import _Criterion


class PointPillarLossMultiClass(_Criterion):
    def __init__(self, args):
        super().__init__(args)
        self.cls_num = args["num_class"]

    def forward(self, output_dict, target_dict, prefix=""):
        return self._call(output_dict, target_dict, prefix, self.cls_num)
