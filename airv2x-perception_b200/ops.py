"""Thin torch-tensor wrappers over the C-ABI (device memory and streams come from torch; compute does not).

All activations are fp32 NHWC tensors; channel slices of wider NHWC buffers are passed as (view, pixel stride).
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvShape, _ptr, call, stream_ptr

c_int = ctypes.c_int


def _cs(t):
    """pixel stride (elements) of an NHWC tensor or channel-slice view."""
    assert t.dim() == 4 and t.stride(3) == 1, "expected NHWC tensor with contiguous channels"
    cs = t.stride(2)
    assert t.stride(1) == t.shape[2] * cs and t.stride(0) == t.shape[1] * t.shape[2] * cs, "non-dense pixel layout"
    return cs


def _shape(n, h, w, cin, cout, k, s):
    return ConvShape(n, h, w, cin, cout, k, s)


def _f32(t):
    assert t.dtype == torch.float32 and t.is_cuda
    return t


# ----------------------------------------------------------------------------- weights
def pack_conv_weight(w, cout_pad=None):
    """OIHW -> ([kk][cout_pad][cin], [kk][cin][cout_pad])"""
    cout, cin, k, _ = w.shape
    cout_pad = cout_pad or cout
    wf = torch.empty(k * k, cout_pad, cin, device=w.device, dtype=torch.float32)
    wd = torch.empty(k * k, cin, cout_pad, device=w.device, dtype=torch.float32)
    call("a2x_pack_conv_weight", _ptr(_f32(w.contiguous())), c_int(cout), c_int(cin), c_int(k), c_int(cout_pad),
         _ptr(wf), _ptr(wd), stream_ptr())
    return wf, wd


def unpack_conv_wgrad(dwp, cout, cin, k, out=None, accumulate=False):
    cout_pad = dwp.shape[1]
    if out is None:
        out = torch.empty(cout, cin, k, k, device=dwp.device, dtype=torch.float32)
    call("a2x_unpack_conv_wgrad", _ptr(dwp), c_int(cout), c_int(cin), c_int(k), c_int(cout_pad), _ptr(out),
         c_int(int(accumulate)), stream_ptr())
    return out


def pack_deconv_weight(w):
    """[cin][cout][s][s] -> ([(ij,co)][ci], [ij][ci][co])"""
    cin, cout, s, _ = w.shape
    wf = torch.empty(s * s * cout, cin, device=w.device, dtype=torch.float32)
    wd = torch.empty(s * s, cin, cout, device=w.device, dtype=torch.float32)
    call("a2x_pack_deconv_weight", _ptr(_f32(w.contiguous())), c_int(cin), c_int(cout), c_int(s), _ptr(wf), _ptr(wd),
         stream_ptr())
    return wf, wd


def unpack_deconv_wgrad(dwp, cin, cout, s, out=None, accumulate=False):
    if out is None:
        out = torch.empty(cin, cout, s, s, device=dwp.device, dtype=torch.float32)
    call("a2x_unpack_deconv_wgrad", _ptr(dwp), c_int(cin), c_int(cout), c_int(s), _ptr(out), c_int(int(accumulate)),
         stream_ptr())
    return out


# ----------------------------------------------------------------------------- conv
def conv2d_fwd(x, wf, k, stride, out=None, scale=None, shift=None, relu=False):
    n, h, w, cin = x.shape
    cout = wf.shape[1]
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    if out is None:
        out = torch.empty(n, ho, wo, cout, device=x.device, dtype=torch.float32)
    sh = _shape(n, h, w, cin, cout, k, stride)
    call("a2x_conv2d_fwd", ctypes.byref(sh), _ptr(_f32(x)), c_int(_cs(x)), _ptr(wf), _ptr(out), c_int(_cs(out)),
         _ptr(scale), _ptr(shift), c_int(int(relu)), stream_ptr())
    return out


def conv2d_dgrad(dy, wd, k, stride, h, w, out=None, accumulate=False):
    n, ho, wo, cout = dy.shape
    cin = wd.shape[1]
    if out is None:
        out = torch.empty(n, h, w, cin, device=dy.device, dtype=torch.float32)
    sh = _shape(n, h, w, cin, cout, k, stride)
    call("a2x_conv2d_dgrad", ctypes.byref(sh), _ptr(_f32(dy)), c_int(_cs(dy)), _ptr(wd), _ptr(out), c_int(_cs(out)),
         c_int(int(accumulate)), stream_ptr())
    return out


def conv2d_wgrad(x, dy, k, stride, out=None):
    """returns packed [kk][cout][cin] (accumulates into `out` if given)"""
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    if out is None:
        out = torch.zeros(k * k, cout, cin, device=x.device, dtype=torch.float32)
    sh = _shape(n, h, w, cin, cout, k, stride)
    call("a2x_conv2d_wgrad", ctypes.byref(sh), _ptr(_f32(x)), c_int(_cs(x)), _ptr(_f32(dy)), c_int(_cs(dy)), _ptr(out),
         stream_ptr())
    return out


def deconv_fwd(x, wf, cout, s, out=None, scale=None, shift=None, relu=False):
    n, h, w, cin = x.shape
    if out is None:
        out = torch.empty(n, h * s, w * s, cout, device=x.device, dtype=torch.float32)
    sh = _shape(n, h, w, cin, cout, s, s)
    call("a2x_deconv_fwd", ctypes.byref(sh), _ptr(_f32(x)), c_int(_cs(x)), _ptr(wf), _ptr(out), c_int(_cs(out)),
         _ptr(scale), _ptr(shift), c_int(int(relu)), stream_ptr())
    return out


def deconv_dgrad(dy, wd, s, out=None, accumulate=False):
    n, h2, w2, cout = dy.shape
    cin = wd.shape[1]
    h, w = h2 // s, w2 // s
    if out is None:
        out = torch.empty(n, h, w, cin, device=dy.device, dtype=torch.float32)
    sh = _shape(n, h, w, cin, cout, s, s)
    call("a2x_deconv_dgrad", ctypes.byref(sh), _ptr(_f32(dy)), c_int(_cs(dy)), _ptr(wd), _ptr(out), c_int(_cs(out)),
         c_int(int(accumulate)), stream_ptr())
    return out


def deconv_wgrad(x, dy, s, out=None):
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    if out is None:
        out = torch.zeros(s * s, cin, cout, device=x.device, dtype=torch.float32)
    sh = _shape(n, h, w, cin, cout, s, s)
    call("a2x_deconv_wgrad", ctypes.byref(sh), _ptr(_f32(x)), c_int(_cs(x)), _ptr(_f32(dy)), c_int(_cs(dy)), _ptr(out),
         stream_ptr())
    return out
