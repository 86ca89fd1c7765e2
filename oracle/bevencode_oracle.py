"""TEST INFRASTRUCTURE ONLY: CPU restatement (torch functional, from a state_dict) of the camera branch's BEV encoder,
the consumer of the Lift-Splat pooling. Groundwork for the next kernels: no CUDA path consumes it yet.
Only tests/ and scripts/ may import it.

Follows:
  opencood/models/sub_modules/lss_submodule.py:312-349  BevEncode (7x7 s2 stem, ResNet-18 layer1-3, Up x4, up2)
  opencood/models/sub_modules/lss_submodule.py:23-47    Up (bilinear align_corners=True, pad, concat [skip, up], 2 x conv3x3+BN+ReLU)
  torchvision.models.resnet.BasicBlock                   conv3x3-BN-ReLU-conv3x3-BN (+ 1x1 s2 conv-BN shortcut) -ReLU
Pinned against the REAL reference module (random weights, eval and train mode) by scripts/make_golden_bevencode.py.
"""
import torch
import torch.nn.functional as F


def _bn(x, sd, pre, training, buffers, eps=1e-5, momentum=0.1):
    """nn.BatchNorm2d defaults (eps 1e-5, momentum 0.1 — NOT the 1e-3 / 0.01 of the LiDAR backbone)"""
    rm, rv = sd[pre + ".running_mean"].clone(), sd[pre + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, sd[pre + ".weight"], sd[pre + ".bias"], training, momentum, eps)
    if training and buffers is not None:
        buffers[pre + ".running_mean"], buffers[pre + ".running_var"] = rm, rv
    return y


def basic_block(sd, pre, x, stride, training, buffers):
    idt = x
    y = F.conv2d(x, sd[pre + ".conv1.weight"], stride=stride, padding=1)
    y = F.relu(_bn(y, sd, pre + ".bn1", training, buffers))
    y = F.conv2d(y, sd[pre + ".conv2.weight"], padding=1)
    y = _bn(y, sd, pre + ".bn2", training, buffers)
    if pre + ".downsample.0.weight" in sd:
        idt = _bn(F.conv2d(x, sd[pre + ".downsample.0.weight"], stride=stride), sd, pre + ".downsample.1", training, buffers)
    return F.relu(y + idt)


def up(sd, pre, x1, x2, scale, training, buffers):
    x1 = F.interpolate(x1, scale_factor=scale, mode="bilinear", align_corners=True)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    x = torch.cat([x2, x1], 1)
    x = F.relu(_bn(F.conv2d(x, sd[pre + ".conv.0.weight"], padding=1), sd, pre + ".conv.1", training, buffers))
    return F.relu(_bn(F.conv2d(x, sd[pre + ".conv.3.weight"], padding=1), sd, pre + ".conv.4", training, buffers))


def bev_encode(sd, x, training=False, buffers=None, pre=""):
    """x [B, inC, H, W] -> [B, outC, H, W] (H, W multiples of 8)"""
    p = pre
    x = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], stride=2, padding=3), sd, p + "bn1", training, buffers))
    x1 = x
    for b in range(2):
        x1 = basic_block(sd, "%slayer1.%d" % (p, b), x1, 1, training, buffers)
    x = x1
    for li, name in ((2, "layer2"), (3, "layer3")):
        for b in range(2):
            x = basic_block(sd, "%s%s.%d" % (p, name, b), x, 2 if b == 0 else 1, training, buffers)
    x = up(sd, p + "up1", x, x1, 4, training, buffers)
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = F.relu(_bn(F.conv2d(x, sd[p + "up2.1.weight"], padding=1), sd, p + "up2.2", training, buffers))
    return F.conv2d(x, sd[p + "up2.4.weight"], sd[p + "up2.4.bias"])
