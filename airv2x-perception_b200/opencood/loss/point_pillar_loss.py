"""`opencood.loss.point_pillar_loss` on the B200 path: `PointPillarLoss(args)` of the legacy `point_pillar_*` models
(loss/point_pillar_loss.py:77-215: one logit per anchor, no objectness) on `a2x_det_loss_legacy`."""
from ...det_loss import _Criterion


class PointPillarLoss(_Criterion):
    legacy = True

    def forward(self, output_dict, target_dict, prefix=""):
        return self._call(output_dict, target_dict, prefix, 1)
