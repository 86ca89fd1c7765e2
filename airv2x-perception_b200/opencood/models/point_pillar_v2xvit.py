"""`opencood.models.point_pillar_v2xvit.PointPillarV2XVit` on the B200 kernels.

Same registry name / class name / constructor, same `hypes_yaml` keys (`pillar_vfe`, `point_pillar_scatter`,
`base_bev_backbone`, `shrink_header`, `compression`, `transformer.encoder.*`, `max_cav`, `anchor_number`, `voxel_size`),
same `state_dict` keys and shapes (13 543 561 parameters for V2XR_v2xvit.yaml), same
`forward(data_dict) -> {"psm","rm","mask","each_mask","comm_rate"}` as opencood/models/point_pillar_v2xvit.py:14-185 of
the reference (input: `data_dict["processed_lidar"]`, `record_len`, `pairwise_t_matrix`; the model itself supplies a
zero prior encoding and an identity spatial correction, :96-105). Parameter containers only; eval forward, the
reference-style training loop and the fused `train_step()` (see point_pillar_cobevt.py); no CPU fallback.
"""
import torch

from ...pplegacy_engine import LegacyV2XViTEngine
from .airv2x_v2xvit import Airv2xV2XVit, _TransformerParams
from .point_pillar_cobevt import _LegacyFusionModel


class PointPillarV2XVit(_LegacyFusionModel):
    ENGINE = LegacyV2XViTEngine

    def __init__(self, args, precision="split3"):
        super().__init__()
        self._init_encoder(args, precision)
        self.fusion_net = _TransformerParams(args["transformer"])
        self._init_heads(args)
        self.discrete_ratio = args["voxel_size"][0]

    _dropouts = Airv2xV2XVit._dropouts               # transformer.encoder.*.dropout, same keys as the airv2x yaml
    _dropout_state = Airv2xV2XVit._dropout_state

    def _forward_train(self, P, lidar, layout, data_dict, drops):
        return self.engine.forward_train(P, lidar, layout, data_dict["pairwise_t_matrix"], drops)

    def forward(self, data_dict):
        if self.training and torch.is_grad_enabled():
            return self._train_forward_autograd(data_dict)
        lidar, layout = self._inputs(data_dict)
        heads, aux = self.engine.forward(self._param_dict(), lidar, layout, self.training,
                                         pairwise=data_dict["pairwise_t_matrix"])
        return self._outputs(heads, aux)
