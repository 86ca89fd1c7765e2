"""Thin torch-tensor wrappers over the C-ABI (device memory and streams come from torch; compute does not).

All activations are fp32 NHWC tensors; channel slices of wider NHWC buffers are passed as (view, pixel stride).
"""
import ctypes
import os

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
class Weights(ctypes.Structure):
    _fields_ = [("w32", ctypes.c_void_p), ("w16", ctypes.c_void_p)]


class PackedW:
    """packed conv / deconv weights: forward (f32, f16) and data-gradient (d32, d16) layouts; *16 are bf16 [2,...]"""

    __slots__ = ("f32", "f16", "d32", "d16", "fwd", "dgrad")

    def __init__(self, f32, f16, d32, d16):
        self.f32, self.f16, self.d32, self.d16 = f32, f16, d32, d16
        self.fwd = Weights(f32.data_ptr(), f16.data_ptr())
        self.dgrad = Weights(d32.data_ptr(), d16.data_ptr())


def pack_conv_weight(w, cout_pad=None, out=None):
    """OIHW -> PackedW with f32 [kk][cout_pad][cin], f16 [2][kk][cout_pad][cin], d32 [kk][cin][cout_pad], d16 [2][...]"""
    cout, cin, k, _ = w.shape
    cout_pad = cout_pad or cout
    if out is None:
        dev = w.device
        out = PackedW(torch.empty(k * k, cout_pad, cin, device=dev), torch.empty(2, k * k, cout_pad, cin, device=dev, dtype=torch.bfloat16),
                      torch.empty(k * k, cin, cout_pad, device=dev), torch.empty(2, k * k, cin, cout_pad, device=dev, dtype=torch.bfloat16))
    call("a2x_pack_conv_weight", _ptr(_f32(w.contiguous())), c_int(cout), c_int(cin), c_int(k), c_int(cout_pad),
         _ptr(out.f32), _ptr(out.f16), _ptr(out.d32), _ptr(out.d16), stream_ptr())
    return out


def unpack_conv_wgrad(dwp, cout, cin, k, out=None, accumulate=False):
    """packed [kk][cout_pad][cin] -> OIHW"""
    cout_pad = dwp.shape[1]
    if out is None:
        out = torch.empty(cout, cin, k, k, device=dwp.device, dtype=torch.float32)
    call("a2x_unpack_conv_wgrad", _ptr(dwp), c_int(cout), c_int(cin), c_int(k), c_int(cout_pad), _ptr(out),
         c_int(int(accumulate)), stream_ptr())
    return out


def pack_deconv_weight(w, out=None):
    """[cin][cout][s][s] -> PackedW with f32 [1][(ij,co)][ci], d32 [ss][ci][co] (+ bf16 pairs)"""
    cin, cout, s, _ = w.shape
    if out is None:
        dev = w.device
        out = PackedW(torch.empty(1, s * s * cout, cin, device=dev), torch.empty(2, 1, s * s * cout, cin, device=dev, dtype=torch.bfloat16),
                      torch.empty(s * s, cin, cout, device=dev), torch.empty(2, s * s, cin, cout, device=dev, dtype=torch.bfloat16))
    call("a2x_pack_deconv_weight", _ptr(_f32(w.contiguous())), c_int(cin), c_int(cout), c_int(s), _ptr(out.f32),
         _ptr(out.f16), _ptr(out.d32), _ptr(out.d16), stream_ptr())
    return out


def unpack_deconv_wgrad(dwp, cin, cout, s, out=None, accumulate=False):
    if out is None:
        out = torch.empty(cin, cout, s, s, device=dwp.device, dtype=torch.float32)
    call("a2x_unpack_deconv_wgrad", _ptr(dwp), c_int(cin), c_int(cout), c_int(s), _ptr(out), c_int(int(accumulate)),
         stream_ptr())
    return out


# ----------------------------------------------------------------------------- batched pack / unpack (one launch)
class PackJob(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("f32", ctypes.c_void_p), ("f16", ctypes.c_void_p), ("d32", ctypes.c_void_p),
                ("d16", ctypes.c_void_p), ("kind", c_int), ("a", c_int), ("b", c_int), ("kk", c_int),
                ("cout_pad", c_int), ("row0", c_int), ("elem_begin", ctypes.c_longlong), ("elems", ctypes.c_longlong)]


class UnpackJob(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("kind", c_int), ("a", c_int), ("b", c_int),
                ("kk", c_int), ("cout_pad", c_int), ("row0", c_int), ("elem_begin", ctypes.c_longlong),
                ("elems", ctypes.c_longlong)]


PACK_F32_PLANES = True   # engines in split (bf16 x 3) mode set False: the fp32 weight planes feed no GEMM there


def conv_pack_job(w, pk, row0=0, f32=None):
    """OIHW parameter `w` -> rows [row0, row0 + cout) of the PackedW `pk` (cout_pad = pk.f32.shape[1]); f32 False: only
    the bf16 planes are written"""
    cout, cin, k, _ = w.shape
    f32 = PACK_F32_PLANES if f32 is None else f32
    return PackJob(w.data_ptr(), pk.f32.data_ptr() if f32 else None, pk.f16.data_ptr(), pk.d32.data_ptr() if f32 else None,
                   pk.d16.data_ptr(), 0, cout, cin, k * k, pk.f32.shape[1], row0, 0, k * k * cout * cin)


def deconv_pack_job(w, pk, f32=None):
    cin, cout, s, _ = w.shape
    f32 = PACK_F32_PLANES if f32 is None else f32
    return PackJob(w.data_ptr(), pk.f32.data_ptr() if f32 else None, pk.f16.data_ptr(), pk.d32.data_ptr() if f32 else None,
                   pk.d16.data_ptr(), 1, cin, cout, s * s, 0, 0, 0, cin * cout * s * s)


def copy_pack_job(src, dst, row0):
    return PackJob(src.data_ptr(), dst.data_ptr(), None, None, None, 2, 0, 0, 0, 0, row0, 0, src.numel())


def conv_unpack_job(dwp, dw, row0=0):
    """packed [kk][cout_pad][cin] rows [row0, row0 + cout) -> OIHW gradient `dw`"""
    cout, cin, k, _ = dw.shape
    return UnpackJob(dwp.data_ptr(), dw.data_ptr(), 0, cout, cin, k * k, dwp.shape[1], row0, 0, dw.numel())


def deconv_unpack_job(dwp, dw):
    cin, cout, s, _ = dw.shape
    return UnpackJob(dwp.data_ptr(), dw.data_ptr(), 1, cin, cout, s * s, 0, 0, 0, dw.numel())


def sums_unpack_job(sums, out, row0=0):
    """float out[i] = (float) double sums[row0 + i]"""
    return UnpackJob(sums.data_ptr(), out.data_ptr(), 2, 0, 0, 0, 0, row0, 0, out.numel())


class JobTable:
    """A device-resident job table (uploaded once per distinct set of pointers)."""

    def __init__(self, jobs, device):
        assert 0 < len(jobs) <= 128, "at most 128 jobs per batched launch"
        pos = 0
        for j in jobs:
            j.elem_begin = pos
            pos += j.elems
        self.total = pos
        self.n = len(jobs)
        arr = (type(jobs[0]) * len(jobs))(*jobs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.dev = host.to(device)
        self.key = tuple((j.src, getattr(j, "dst", None) or getattr(j, "f32", None)) for j in jobs)


def pack_weights_batched(table):
    call("a2x_pack_weights_batched", _ptr(table.dev), c_int(table.n), c_ll(table.total), stream_ptr())


def unpack_wgrads_batched(table):
    call("a2x_unpack_wgrads_batched", _ptr(table.dev), c_int(table.n), c_ll(table.total), stream_ptr())


# ============================================================================= split-plane activations
class Operand(ctypes.Structure):
    _fields_ = [("hi", ctypes.c_void_p), ("b16", ctypes.c_void_p), ("b16_plane", ctypes.c_longlong), ("cs", c_int)]


class Act:
    """An NHWC activation as GEMM operand: `hi` = the fp32 value and, in split mode, `b16`: a bf16 tensor
    [2, n, h, w, c] with plane 0 = h16 = bf16(v), plane 1 = l16 = bf16(v - h16) (the bf16 x 3 GEMM operands)."""

    __slots__ = ("hi", "b16")

    def __init__(self, hi, b16=None):
        self.hi, self.b16 = hi, b16

    @staticmethod
    def empty(shape, device, split):
        hi = torch.empty(tuple(shape), device=device, dtype=torch.float32)
        b16 = torch.empty((2,) + tuple(shape), device=device, dtype=torch.bfloat16) if split else None
        return Act(hi, b16)

    def value(self):
        return self.hi

    def slice_c(self, c0, c1):
        return Act(self.hi[..., c0:c1], None if self.b16 is None else self.b16[..., c0:c1])

    def narrow_n(self, n0, n):
        return Act(self.hi[n0:n0 + n], None if self.b16 is None else self.b16[:, n0:n0 + n])

    @property
    def shape(self):
        return self.hi.shape

    def operand(self):
        cs = _cs(self.hi)
        if self.b16 is None:
            return Operand(self.hi.data_ptr(), None, 0, cs)
        assert _cs(self.b16[0]) == cs
        return Operand(self.hi.data_ptr(), self.b16.data_ptr(), self.b16.stride(0), cs)


def _op(a):
    return ctypes.byref(a.operand())


POISON_SKIPPED_HI = bool(int(os.environ.get("A2X_POISON_HI", "0")))   # debug: NaN-fill fp32 planes that are not written


def _op_planes(a, write_hi):
    """output operand; write_hi False (split mode only): the kernel leaves the fp32 plane unwritten (hi pointer null)"""
    o = a.operand()
    if not write_hi and a.b16 is not None:
        if POISON_SKIPPED_HI:
            a.hi.fill_(float("nan"))
        o.hi = None
    return ctypes.byref(o)


def split(x):
    """fp32 NHWC tensor -> split Act"""
    x = x.contiguous()
    out = Act.empty(x.shape, x.device, True)
    o = out.operand()
    call("a2x_split", _ptr(x), ctypes.c_longlong(x.numel()), ctypes.byref(o), stream_ptr())
    return out


split_tf32 = split


# split-aware conv family ------------------------------------------------------------------------------------
def conv_fwd(x, w, k, stride, out, scale=None, shift=None, relu=False, stats=None, accumulate=False, write_hi=True):
    """x, out: Act; w: PackedW. relu: False/0 none, True/1 ReLU, 2 GELU(erf). accumulate: out += result (residual).
    write_hi False: only the split planes of `out` are stored (every consumer is a split GEMM)."""
    n, h, ww, cin = x.shape
    cout = w.f32.shape[1]
    sh = _shape(n, h, ww, cin, cout, k, stride)
    call("a2x_conv2d_fwd_ex", ctypes.byref(sh), _op(x), ctypes.byref(w.fwd), _op_planes(out, write_hi or accumulate),
         _ptr(scale), _ptr(shift), c_int(int(relu)), c_int(int(accumulate)), _ptr(stats), stream_ptr())
    return out


def linear_fwd(x, w, out, bias=None, act=0, accumulate=False):
    """token-wise nn.Linear on an NHWC token tensor = 1x1 tap-GEMM: out = act(x W^T + bias (+ out))"""
    return conv_fwd(x, w, 1, 1, out, shift=bias, relu=act, accumulate=accumulate)


def linear_dropout_residual_fwd(x, w, y, bias=None, residual=None, drop=None, site=0, elem_offset=0):
    """y (dense fp32 tensor) = residual + dropout(x W^T + bias) in the GEMM epilogue; the mask is the one
    dropout_mask(.., drop, site) exports, element elem_offset + i for y's element i (y a slice of the site's tensor).
    residual may be y itself; drop None: no dropout."""
    n, h, ww, cin = x.shape
    cout = w.f32.shape[1]
    assert y.is_contiguous() and y.shape[3] == cout and (residual is None or residual.shape == y.shape)
    sh = _shape(n, h, ww, cin, cout, 1, 1)
    call("a2x_linear_dropout_residual_fwd", ctypes.byref(sh), _op(x), ctypes.byref(w.fwd), _ptr(y), c_int(cout), _ptr(bias),
         _ptr(residual), ctypes.c_ulonglong(drop.seed if drop is not None else 0), ctypes.c_uint(site),
         c_f(drop.p if drop is not None else 0.0), c_ll(elem_offset), stream_ptr())
    return y


# transformer fusion token kernels --------------------------------------------------------------------------------
def _rows(t):
    return t.shape[0] * t.shape[1] * t.shape[2]


def layernorm_fwd(x, gamma, beta, out, eps=1e-5):
    """x: NHWC tensor; out: Act"""
    call("a2x_layernorm_fwd", _ptr(x), c_int(_cs(x)), c_ll(_rows(x)), c_int(x.shape[3]), _ptr(gamma), _ptr(beta), c_f(eps),
         _op(out), stream_ptr())
    return out


def agent_mean_layernorm(x, B, L, gamma, beta, out, eps=1e-5):
    """x: dense [B*L, H, W, C]; out: Act [B, H, W, C]"""
    assert x.is_contiguous() and x.shape[0] == B * L
    call("a2x_agent_mean_layernorm", _ptr(x), c_int(B), c_int(L), c_ll(x.shape[1] * x.shape[2]), c_int(x.shape[3]),
         _ptr(gamma), _ptr(beta), c_f(eps), _op(out), stream_ptr())
    return out


def regroup(src, scene_start, scene_len, B, L, out):
    """src: dense [N, H, W, C]; out: Act [B*L, H, W, C] (zero padded agents)"""
    assert src.is_contiguous() and out.hi.is_contiguous()
    call("a2x_regroup", _ptr(src), _ptr(scene_start), _ptr(scene_len), c_int(B), c_int(L),
         c_ll(src.shape[1] * src.shape[2] * src.shape[3]), _op(out), stream_ptr())
    return out


def regroup_ptrs(ptr_table, img_shape, scene_start, scene_len, B, L, out):
    """ptr_table: int64 device tensor [N] of per-agent image pointers (possibly peer-GPU memory); img_shape: (H, W, C)"""
    h, w, c = img_shape
    call("a2x_regroup_ptrs", _ptr(ptr_table), _ptr(scene_start), _ptr(scene_len), c_int(B), c_int(L), c_ll(h * w * c),
         _op(out), stream_ptr())
    return out


def window_attention_fwd(qkv, bias_table, key_mask, B, L, heads, dim_head, window, grid_mode, out):
    """qkv: dense [B*L, H, W, 3*heads*dim_head]; out: Act [B*L, H, W, heads*dim_head]"""
    assert qkv.is_contiguous() and out.hi.is_contiguous()
    _, H, W, _ = qkv.shape
    call("a2x_window_attention_fwd", _ptr(qkv), _ptr(bias_table), _ptr(key_mask), c_int(B), c_int(L), c_int(H), c_int(W),
         c_int(heads), c_int(dim_head), c_int(window), c_int(int(grid_mode)), c_f(dim_head ** -0.5), _op(out),
         stream_ptr())
    return out


def warp_affine_fwd(src, theta, out, align_corners=False, nearest=False):
    """src: NHWC [n, hi, wi, c]; theta: [n, 2, 3] f32; out: Act [n, ho, wo, c]"""
    n, hi, wi, c = src.shape
    _, ho, wo, _ = out.shape
    call("a2x_warp_affine_fwd", _ptr(src), c_int(_cs(src)), _ptr(theta.contiguous()), c_int(n), c_int(hi), c_int(wi),
         c_int(c), c_int(ho), c_int(wo), c_int(int(align_corners)), c_int(int(nearest)), _op(out), stream_ptr())
    return out


def warp_affine_bwd(dout, theta, dsrc, align_corners=False):
    """dsrc (zeroed by the caller) += bilinear^T dout"""
    n, hi, wi, c = dsrc.shape
    _, ho, wo, _ = dout.shape
    call("a2x_warp_affine_bwd", _ptr(dout), c_int(_cs(dout)), _ptr(theta.contiguous()), c_int(n), c_int(hi), c_int(wi),
         c_int(c), c_int(ho), c_int(wo), c_int(int(align_corners)), _ptr(dsrc), c_int(_cs(dsrc)), stream_ptr())
    return dsrc


BN_BWD_REPLICAS = 8  # A2X_BN_BWD_REPLICAS: copies of the [2C] BN-backward accumulators


def bn_bwd_sums_len(C):
    """A2X_BN_BWD_SUMS(C): doubles to allocate (and zero) for the BN-backward accumulators"""
    return 2 * C * BN_BWD_REPLICAS + 1


class BnBwdStats(ctypes.Structure):
    _fields_ = [("z", ctypes.c_void_p), ("z_cs", c_int), ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
                ("mean", ctypes.c_void_p), ("invstd", ctypes.c_void_p), ("sums", ctypes.c_void_p)]


def conv_dgrad(dy, w, k, stride, dx, accumulate=False, bn_stats=None):
    """dy: Act; dx: plain NHWC tensor [n,h,w,cin]. bn_stats = (z, scale, shift, mean, invstd, sums) of the layer that
    consumes dx as its dy: pass 1 of its BN+ReLU backward is accumulated in the epilogue (3x3 stride 1)."""
    n, h, ww, cin = dx.shape
    cout = dy.shape[3]
    sh = _shape(n, h, ww, cin, cout, k, stride)
    st = None
    if bn_stats is not None:
        z, scale, shift, mean, invstd, sums = bn_stats
        st = ctypes.byref(BnBwdStats(z.data_ptr(), _cs(z), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                     invstd.data_ptr(), sums.data_ptr()))
    call("a2x_conv2d_dgrad_ex", ctypes.byref(sh), _op(dy), ctypes.byref(w.dgrad), _ptr(dx), c_int(_cs(dx)),
         c_int(int(accumulate)), st, stream_ptr())
    return dx


def conv_wgrad(x, dy, k, stride, dwp):
    """x, dy: Act (split only if both are); dwp: zeroed packed [kk][cout][cin]"""
    n, h, ww, cin = x.shape
    cout = dy.shape[3]
    sh = _shape(n, h, ww, cin, cout, k, stride)
    call("a2x_conv2d_wgrad", ctypes.byref(sh), _op(x), _op(dy), _ptr(dwp), stream_ptr())
    return dwp


def deconv_fwd(x, w, cout, s, out, scale=None, shift=None, relu=False, stats=None, write_hi=True):
    n, h, ww, cin = x.shape
    sh = _shape(n, h, ww, cin, cout, s, s)
    call("a2x_deconv_fwd", ctypes.byref(sh), _op(x), ctypes.byref(w.fwd), _op_planes(out, write_hi), _ptr(scale), _ptr(shift),
         c_int(int(relu)), _ptr(stats), stream_ptr())
    return out


def deconv_dgrad(dy, w, s, dx, accumulate=False):
    n, h, ww, cin = dx.shape
    cout = dy.shape[3]
    sh = _shape(n, h, ww, cin, cout, s, s)
    call("a2x_deconv_dgrad", ctypes.byref(sh), _op(dy), ctypes.byref(w.dgrad), _ptr(dx), c_int(_cs(dx)),
         c_int(int(accumulate)), stream_ptr())
    return dx


def deconv_wgrad(x, dy, s, dwp):
    n, h, ww, cin = x.shape
    cout = dy.shape[3]
    sh = _shape(n, h, ww, cin, cout, s, s)
    call("a2x_deconv_wgrad", ctypes.byref(sh), _op(x), _op(dy), _ptr(dwp), stream_ptr())
    return dwp


# BN / elementwise ---------------------------------------------------------------------------------------------
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_d = ctypes.c_double


def _npix(t):
    return t.shape[0] * t.shape[1] * t.shape[2]


def channel_stats(x, sums):
    call("a2x_channel_stats", _ptr(x), c_int(_cs(x)), c_ll(_npix(x)), c_int(x.shape[3]), _ptr(sums), stream_ptr())


def bn_finalize(sums, count, gamma, beta, n_updates, rm, rv, scale, shift, mean, invstd, eps=1e-3, momentum=0.01):
    call("a2x_bn_finalize", _ptr(sums), c_d(float(count)), _ptr(gamma), _ptr(beta), c_f(eps), c_f(momentum),
         c_int(n_updates), _ptr(rm), _ptr(rv), c_int(scale.numel()), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd),
         stream_ptr())


def bn_eval_affine(gamma, beta, rm, rv, scale, shift, eps=1e-3):
    call("a2x_bn_eval_affine", _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), c_f(eps), c_int(scale.numel()), _ptr(scale),
         _ptr(shift), stream_ptr())


def bn_train_act(z, sums, count, gamma, beta, n_updates, rm, rv, scale, shift, mean, invstd, relu, out, eps=1e-3,
                 momentum=0.01, write_hi=True):
    """fused bn_finalize + affine_act (train mode): out = relu?(BN_batch(z)); scale/shift/mean/invstd are published.
    write_hi False: only the bf16 split planes of `out` are written (its consumers are split GEMMs)."""
    call("a2x_bn_train_act", _ptr(z), c_int(_cs(z)), _ptr(sums), c_d(float(count)), _ptr(gamma), _ptr(beta), c_f(eps),
         c_f(momentum), c_int(n_updates), _ptr(rm), _ptr(rv), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd),
         c_int(int(relu)), _op_planes(out, write_hi), c_ll(_npix(z)), c_int(z.shape[3]), stream_ptr())
    return out


def affine_act(x, scale, shift, relu, out, mask=None, write_hi=True):
    """x: NHWC tensor; out: Act (write_hi False: split planes only)"""
    call("a2x_affine_act", _ptr(x), c_int(_cs(x)), _ptr(scale), _ptr(shift), c_int(int(relu)), _ptr(mask),
         _op_planes(out, write_hi),
         c_ll(_npix(x)), c_int(x.shape[3]), stream_ptr())
    return out


def bn_relu_bwd(dy, z, scale, shift, mean, invstd, sums, dz, dgamma, dbeta, accumulate=False, sums_ready=False,
                write_hi=True):
    """dy, z: NHWC tensors; dz: Act; sums: zeroed [2C] double (already filled by a fused dgrad epilogue if sums_ready)"""
    npix, C = _npix(dy), dy.shape[3]
    if not sums_ready:
        call("a2x_bn_relu_bwd_reduce", _ptr(dy), c_int(_cs(dy)), _ptr(z), c_int(_cs(z)), _ptr(scale), _ptr(shift),
             _ptr(mean), _ptr(invstd), c_ll(npix), c_int(C), _ptr(sums), stream_ptr())
    call("a2x_bn_relu_bwd_apply", _ptr(dy), c_int(_cs(dy)), _ptr(z), c_int(_cs(z)), _ptr(scale), _ptr(shift), _ptr(mean),
         _ptr(invstd), _ptr(sums), c_d(float(npix)), _op_planes(dz, write_hi), c_ll(npix), c_int(C),
         _ptr(dgamma), _ptr(dbeta), c_int(int(accumulate)), stream_ptr())
    return dz


def relu_bwd(dy, y, out, mask=None):
    """g = dy * (y > 0) * mask ; dy, y NHWC tensors (y may be None); out: Act"""
    call("a2x_relu_bwd", _ptr(dy), c_int(_cs(dy)), _ptr(y), c_int(_cs(y) if y is not None else 0), _ptr(mask),
         _op(out), c_ll(_npix(dy)), c_int(dy.shape[3]), stream_ptr())
    return out


def sums_to_float(sums, C, out, accumulate=False):
    call("a2x_sums_to_float", _ptr(sums), c_int(C), _ptr(out), c_int(int(accumulate)), stream_ptr())


def count_nonzero(x, out_u64):
    call("a2x_count_nonzero", _ptr(x), c_ll(x.numel()), _ptr(out_u64), stream_ptr())


# PFN ------------------------------------------------------------------------------------------------------------
class PfnGeom(ctypes.Structure):
    _fields_ = [("voxel_x", c_f), ("voxel_y", c_f), ("voxel_z", c_f), ("x_offset", c_f), ("y_offset", c_f),
                ("z_offset", c_f), ("nx", c_int), ("ny", c_int)]


def pfn_geom(voxel_size, lidar_range, nx, ny):
    """airv2x_pillar_vfe.py:84-89 (offsets computed in double like the reference, then cast to fp32)."""
    vx, vy, vz = [float(v) for v in voxel_size]
    return PfnGeom(vx, vy, vz, vx / 2 + lidar_range[0], vy / 2 + lidar_range[1], vz / 2 + lidar_range[2], nx, ny)


class PfnSegments(ctypes.Structure):
    _fields_ = [("seg_ids", ctypes.c_void_p), ("seg_counts", ctypes.c_void_p), ("seg_cap", c_int), ("nseg", c_int)]


def pfn_segments(seg_ids, seg_counts, cap):
    """seg_ids: int32 device tensor (slab index per segment); seg_counts: int32 device tensor indexed by slab"""
    return PfnSegments(seg_ids.data_ptr(), seg_counts.data_ptr(), cap, seg_ids.numel())


def _seg(seg):
    return ctypes.byref(seg) if seg is not None else ctypes.c_void_p(0)


def _m(vox, seg):
    return 0 if seg is not None else vox.shape[0]


def pfn_moments(vox, num, coords, geom, moments, seg=None):
    call("a2x_pfn_moments", _ptr(vox), _ptr(num), _ptr(coords), c_ll(_m(vox, seg)), ctypes.byref(geom), _seg(seg),
         _ptr(moments), stream_ptr())


def pfn_stats_finalize(moments, rows, w, gamma, beta, n_updates, rm, rv, scale, shift, mean, invstd, eps=1e-3,
                       momentum=0.01, seg=None):
    call("a2x_pfn_stats_finalize", _ptr(moments), c_d(float(rows)), _seg(seg), _ptr(w), _ptr(gamma), _ptr(beta), c_f(eps),
         c_f(momentum), c_int(n_updates), _ptr(rm), _ptr(rv), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd),
         stream_ptr())


def pfn_scatter(vox, num, coords, geom, w, scale, shift, agent_map, canvas, pillar_out=None, amax=None, seg=None, nz=None,
                write_hi=True):
    """nz: int64 device counter accumulating count_nonzero of what is scattered (caller zeroes it); write_hi False (split
    canvas): only the bf16 planes are written"""
    call("a2x_pfn_scatter_ex", _ptr(vox), _ptr(num), _ptr(coords), c_ll(_m(vox, seg)), ctypes.byref(geom), _seg(seg), _ptr(w),
         _ptr(scale), _ptr(shift), _ptr(agent_map), _op_planes(canvas, write_hi), _ptr(pillar_out), _ptr(amax), _ptr(nz),
         stream_ptr())


def pfn_bwd(vox, num, coords, geom, w, scale, shift, mean, invstd, agent_map, dcanvas, amax, moments, rows, acc_ws, dw,
            dgamma, dbeta, accumulate=False, seg=None):
    call("a2x_pfn_bwd", _ptr(vox), _ptr(num), _ptr(coords), c_ll(_m(vox, seg)), ctypes.byref(geom), _seg(seg), _ptr(w),
         _ptr(scale),
         _ptr(shift), _ptr(mean), _ptr(invstd), _ptr(agent_map), _ptr(dcanvas), _ptr(amax), _ptr(moments),
         c_d(float(rows)), _ptr(acc_ws), _ptr(dw), _ptr(dgamma), _ptr(dbeta), c_int(int(accumulate)), stream_ptr())


# communication / fusion -----------------------------------------------------------------------------------------
def comm_confidence(psm, ncls, conf):
    call("a2x_comm_confidence", _ptr(psm), c_int(_cs(psm)), c_int(ncls), c_ll(_npix(psm)), _ptr(conf), stream_ptr())


def comm_smooth_mask(conf, gw, gb, ksz, n, h, w, thr, write_mask, smooth, mask):
    call("a2x_comm_smooth_mask", _ptr(conf), _ptr(gw), _ptr(gb), c_int(ksz), c_int(n), c_int(h), c_int(w), c_f(thr),
         c_int(int(write_mask)), _ptr(smooth), _ptr(mask), stream_ptr())


def comm_topk_mask(smooth, n, hw, k_dev, mask):
    call("a2x_comm_topk_mask", _ptr(smooth), c_int(n), c_int(hw), _ptr(k_dev), _ptr(mask), stream_ptr())


def comm_rate_ego(mask, hw, n_scenes, scene_start, scene_len, ones):
    call("a2x_comm_rate_ego", _ptr(mask), c_int(hw), c_int(n_scenes), _ptr(scene_start), _ptr(scene_len), _ptr(ones),
         stream_ptr())


def resize_bilinear(src, dst):
    """src: [n, h, w] float; dst: [n, H, W] (F.interpolate bilinear, align_corners=False)"""
    call("a2x_resize_bilinear", _ptr(src), c_int(src.shape[0]), c_int(src.shape[1]), c_int(src.shape[2]), _ptr(dst),
         c_int(dst.shape[1]), c_int(dst.shape[2]), stream_ptr())
    return dst


def mask_compact(x, mask, force_all, hdr, idx, vals):
    """x: NHWC [1, h, w, C]; mask: [h, w] float; hdr: int32 [>=2]; idx: int32 [hw]; vals: float [hw, C]"""
    _, h, w, c = x.shape
    call("a2x_mask_compact", _ptr(x), c_int(_cs(x)), _ptr(mask), c_int(int(force_all)), c_int(h * w), c_int(c), _ptr(hdr),
         _ptr(idx), _ptr(vals), stream_ptr())


def mask_decompact_ptrs(ptr_table, off_idx_bytes, off_vals_bytes, n_agents, dst):
    """ptr_table: int64 device tensor [n] of record-buffer base pointers; dst: dense [n, h, w, C] (zero-filled here)"""
    n, h, w, c = dst.shape
    assert n == n_agents and dst.is_contiguous()
    call("a2x_mask_decompact_ptrs", _ptr(ptr_table), c_ll(off_idx_bytes), c_ll(off_vals_bytes), c_int(n), c_int(h * w),
         c_int(c), _ptr(dst), stream_ptr())
    return dst


def att_fuse_fwd(x, out):
    """x: dense NHWC tensor [n_agents,h,w,c] of one scene; out: Act [1,h,w,c] (or [h,w,c])"""
    n, h, w, c = x.shape
    call("a2x_att_fuse_fwd", _ptr(x), c_int(n), c_int(h * w), c_int(c), _op(out), stream_ptr())


def att_fuse_bwd(x, dout, dx):
    n, h, w, c = x.shape
    call("a2x_att_fuse_bwd", _ptr(x), _ptr(dout), c_int(n), c_int(h * w), c_int(c), _ptr(dx), stream_ptr())


def det_loss(heads, A, K, targets, pos, class_ids, cls_weight, reg_coe, npos_ws, dheads, loss3, legacy=False):
    B, H, W, cs = heads.shape[0], heads.shape[1], heads.shape[2], _cs(heads)
    if legacy:   # PointPillarLoss of the legacy point_pillar_* models (K = 1, no objectness)
        call("a2x_det_loss_legacy", _ptr(heads), c_int(cs), c_int(B), c_ll(H * W), c_int(A), _ptr(targets), _ptr(pos),
             c_f(cls_weight), c_f(reg_coe), _ptr(npos_ws), _ptr(dheads), c_int(_cs(dheads) if dheads is not None else 0),
             _ptr(loss3), stream_ptr())
        return
    call("a2x_det_loss", _ptr(heads), c_int(cs), c_int(B), c_ll(H * W), c_int(A), c_int(K), _ptr(targets), _ptr(pos),
         _ptr(class_ids), c_f(cls_weight), c_f(reg_coe), _ptr(npos_ws), _ptr(dheads),
         c_int(_cs(dheads) if dheads is not None else 0), _ptr(loss3), stream_ptr())


# voxelisation ---------------------------------------------------------------------------------------------------
def voxelize_workspace_bytes(n_agents, total_points, nx, ny, nz, cap):
    lib = _lib.load()
    lib.a2x_voxelize_workspace_bytes.restype = ctypes.c_size_t
    return lib.a2x_voxelize_workspace_bytes(c_int(n_agents), c_ll(total_points), c_int(nx), c_int(ny), c_int(nz), c_int(cap))


def voxelize(points, offsets_dev, n_agents, lidar_range, voxel_size, max_points, max_voxels, cap, workspace, voxels,
             coords, num_points, counts, ego_flags=None, strict_range=False, transforms=None):
    """points: [total,4] f32 device; offsets_dev: int32 [n_agents+1] device; transforms: optional [n_agents,4,4] f32 device
    (agent -> ego projection applied before the range test). Slab outputs (see include/airv2x_b200.h)."""
    rng = (c_f * 6)(*[float(v) for v in lidar_range])
    vs = (c_f * 3)(*[float(v) for v in voxel_size])
    if transforms is not None:
        assert transforms.shape == (n_agents, 4, 4) and transforms.dtype == torch.float32 and transforms.is_contiguous()
    call("a2x_voxelize_ex", _ptr(points), _ptr(offsets_dev), _ptr(transforms), c_int(n_agents), c_ll(points.shape[0]), rng, vs,
         c_int(max_points), c_int(max_voxels), c_int(cap), _ptr(ego_flags), c_int(int(strict_range)), _ptr(workspace),
         ctypes.c_size_t(workspace.numel()),
         _ptr(voxels), _ptr(coords), _ptr(num_points), _ptr(counts), stream_ptr())


def roi_mask(theta, valid, n, h, w, out, align_corners=True):
    call("a2x_roi_mask", _ptr(theta.contiguous()), _ptr(valid), c_int(n), c_int(h), c_int(w), c_int(int(align_corners)),
         _ptr(out), stream_ptr())
    return out


# V2X-ViT fusion ---------------------------------------------------------------------------------------------------
def rte_add(x, emb_table, emb_idx, lin_w, lin_b, vec_ws):
    """x: dense [n, h, w, C] (+= per-agent vector Linear(emb[idx]))"""
    n, h, w, c = x.shape
    assert x.is_contiguous()
    call("a2x_rte_add", _ptr(x), c_int(n), c_ll(h * w), c_int(c), _ptr(emb_table), _ptr(emb_idx), _ptr(lin_w), _ptr(lin_b),
         _ptr(vec_ws), stream_ptr())


def _ptr2(a, b):
    return (ctypes.c_void_p * 2)(a.data_ptr(), b.data_ptr())


def hgt_fold(qw, qb, kw, kb, vw, vb, rel_att, rel_msg, heads, w_fold, b_fold):
    """qw..vb: pairs of tensors (agent type 0, 1); w_fold: [2, 5C, C]; b_fold: [2, 5C]"""
    C = qw[0].shape[1]
    call("a2x_hgt_fold", _ptr2(*qw), _ptr2(*qb), _ptr2(*kw), _ptr2(*kb), _ptr2(*vw), _ptr2(*vb), _ptr(rel_att),
         _ptr(rel_msg), c_int(C), c_int(heads), _ptr(w_fold), _ptr(b_fold), stream_ptr())


def hgt_attention_fwd(qkv, types, key_mask, heads, dim_head, out):
    """qkv: dense [n, h, w, 5C]; types: int32 [n]; key_mask: float [n, h, w]; out: Act [n, h, w, C]"""
    n, h, w, _ = qkv.shape
    assert qkv.is_contiguous() and key_mask.is_contiguous()
    call("a2x_hgt_attention_fwd", _ptr(qkv), _ptr(types), _ptr(key_mask), c_int(n), c_ll(h * w), c_int(heads),
         c_int(dim_head), c_f(dim_head ** -0.5), _op(out), stream_ptr())
    return out


SPLIT_ATTN_CHUNKS = 64   # A2X_SPLIT_ATTN_CHUNKS
_split_partials = {}


def split_attn_fuse(w0, w1, w2, fc1, ln_g, ln_b, fc2, sums_ws, weights_ws, x):
    """x (dense [n, h, w, C]) += split-attention mix of the three window branches; sums_ws [n, C] receives the pooled sums.
    The per-chunk scratch of the deterministic pooling is cached here per (device, n, C)."""
    n, h, w, c = x.shape
    key = (x.device, n, c)
    part = _split_partials.get(key)
    if part is None:
        part = _split_partials[key] = torch.empty(n * SPLIT_ATTN_CHUNKS * c, device=x.device, dtype=torch.float32)
    call("a2x_split_attn_fuse", _ptr(w0), _ptr(w1), _ptr(w2), c_int(n), c_ll(h * w), c_int(c), _ptr(fc1), _ptr(ln_g),
         _ptr(ln_b), _ptr(fc2), _ptr(sums_ws), _ptr(part), _ptr(weights_ws), _ptr(x), stream_ptr())


# transformer fusion backward ----------------------------------------------------------------------------------------
def layernorm_bwd(x, dy, gamma, dx_accum, dgamma, dbeta, eps=1e-5):
    """x, dy, dx_accum: NHWC tensors (dx_accum += LN backward); dgamma/dbeta: zeroed float64 [C]"""
    call("a2x_layernorm_bwd", _ptr(x), c_int(_cs(x)), _ptr(dy), c_int(_cs(dy)), c_ll(_rows(x)), c_int(x.shape[3]),
         _ptr(gamma), c_f(eps), _ptr(dx_accum), c_int(_cs(dx_accum)), _ptr(dgamma), _ptr(dbeta), stream_ptr())


def gelu_fwd(x, out):
    call("a2x_gelu_fwd", _ptr(x), c_ll(x.numel()), _op(out), stream_ptr())
    return out


class Dropout:
    """one training step's dropout state: rate p, the step's 64-bit seed, and a running site counter (every nn.Dropout
    call site of the forward takes the next id; the backward looks the id up again)"""

    def __init__(self, p, seed):
        self.p, self.seed, self.n_sites = float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, 0

    def site(self):
        self.n_sites += 1
        return self.n_sites - 1


def dropout_apply(y, drop, site, out, residual=None, write_hi=True):
    """out (Act) = residual + y * keep / (1 - p); y / residual dense fp32 tensors; drop None or p == 0: plain copy / add"""
    p = drop.p if drop is not None else 0.0
    call("a2x_dropout_apply", _ptr(y), _ptr(residual), c_ll(y.numel()), ctypes.c_ulonglong(drop.seed if drop else 0),
         ctypes.c_uint(site), c_f(p), _op_planes(out, write_hi), stream_ptr())
    return out


def gelu_dropout_fwd(x, drop, site, out):
    call("a2x_gelu_dropout_fwd", _ptr(x), c_ll(x.numel()), ctypes.c_ulonglong(drop.seed), ctypes.c_uint(site), c_f(drop.p),
         _op(out), stream_ptr())
    return out


def gelu_dropout_bwd(dy, x, drop, site, out):
    call("a2x_gelu_dropout_bwd", _ptr(dy), _ptr(x), c_ll(x.numel()), ctypes.c_ulonglong(drop.seed), ctypes.c_uint(site),
         c_f(drop.p), _op(out), stream_ptr())
    return out


def dropout_mask(n, drop, site):
    """uint8 keep flags of a site (test hook: identical masks for the oracle)"""
    m = torch.empty(n, dtype=torch.uint8, device="cuda")
    call("a2x_dropout_mask", c_ll(n), ctypes.c_ulonglong(drop.seed), ctypes.c_uint(site), c_f(drop.p), _ptr(m), stream_ptr())
    return m


def gelu_bwd(dy, x, out):
    call("a2x_gelu_bwd", _ptr(dy), _ptr(x), c_ll(x.numel()), _op(out), stream_ptr())
    return out


def window_attention_bwd(qkv, dout, bias_table, key_mask, B, L, heads, dim_head, window, grid_mode, dqkv, dbias,
                         write_hi=True):
    """dqkv: fp32 tensor [B*L, H, W, 3D], or an Act (split planes written directly; write_hi False skips its fp32 plane)"""
    assert qkv.is_contiguous() and dout.is_contiguous()
    _, H, W, _ = qkv.shape
    if isinstance(dqkv, Act):
        assert dqkv.hi.is_contiguous()
        call("a2x_window_attention_bwd_split", _ptr(qkv), _ptr(dout), _ptr(bias_table), _ptr(key_mask), c_int(B), c_int(L),
             c_int(H), c_int(W), c_int(heads), c_int(dim_head), c_int(window), c_int(int(grid_mode)), c_f(dim_head ** -0.5),
             _op_planes(dqkv, write_hi), _ptr(dbias), stream_ptr())
        return
    assert dqkv.is_contiguous()
    call("a2x_window_attention_bwd", _ptr(qkv), _ptr(dout), _ptr(bias_table), _ptr(key_mask), c_int(B), c_int(L), c_int(H),
         c_int(W), c_int(heads), c_int(dim_head), c_int(window), c_int(int(grid_mode)), c_f(dim_head ** -0.5), _ptr(dqkv),
         _ptr(dbias), stream_ptr())


def agent_mean(x, B, L, out):
    """x: dense [B*L, H, W, C] -> out dense [B, H, W, C]"""
    call("a2x_agent_mean", _ptr(x), c_int(B), c_int(L), c_ll(x.shape[1] * x.shape[2] * x.shape[3]), _ptr(out), stream_ptr())
    return out


def agent_broadcast(src, B, L, scale, dst):
    """dst [B*L, H, W, C] = scale * src [B, H, W, C] repeated over the L agents"""
    call("a2x_agent_broadcast", _ptr(src), c_int(B), c_int(L), c_ll(src.shape[1] * src.shape[2] * src.shape[3]), c_f(scale),
         _ptr(dst), stream_ptr())
    return dst


# V2X-ViT fusion backward ----------------------------------------------------------------------------------------------
def hgt_attention_bwd(qkv, types, key_mask, dout, heads, dim_head, dqkv):
    n, h, w, _ = qkv.shape
    assert qkv.is_contiguous() and dout.is_contiguous() and dqkv.is_contiguous() and key_mask.is_contiguous()
    call("a2x_hgt_attention_bwd", _ptr(qkv), _ptr(types), _ptr(key_mask), _ptr(dout), c_int(n), c_ll(h * w), c_int(heads),
         c_int(dim_head), c_f(dim_head ** -0.5), _ptr(dqkv), stream_ptr())
    return dqkv


def hgt_fold_bwd(dw_fold, db_fold, kw, kb, vw, vb, rel_att, rel_msg, heads, dqw, dqb, dkw, dkb, dvw, dvb, drel_att, drel_msg):
    """dw_fold: [2, 5C, C]; db_fold: [2, 5C]; kw..vb and dqw..dvb: pairs of tensors (agent type 0, 1)"""
    C = kw[0].shape[1]
    call("a2x_hgt_fold_bwd", _ptr(dw_fold), _ptr(db_fold), _ptr2(*kw), _ptr2(*kb), _ptr2(*vw), _ptr2(*vb), _ptr(rel_att),
         _ptr(rel_msg), c_int(C), c_int(heads), _ptr2(*dqw), _ptr2(*dqb), _ptr2(*dkw), _ptr2(*dkb), _ptr2(*dvw), _ptr2(*dvb),
         _ptr(drel_att), _ptr(drel_msg), stream_ptr())


def split_attn_bwd(dx, w0, w1, w2, fc1, ln_g, ln_b, fc2, sums_saved, weights_saved, dw_ws, dgap_ws, d0, d1, d2, dfc1, dln_g,
                   dln_b, dfc2):
    n, h, w, c = dx.shape
    call("a2x_split_attn_bwd", _ptr(dx), _ptr(w0), _ptr(w1), _ptr(w2), c_int(n), c_ll(h * w), c_int(c), _ptr(fc1), _ptr(ln_g),
         _ptr(ln_b), _ptr(fc2), _ptr(sums_saved), _ptr(weights_saved), _ptr(dw_ws), _ptr(dgap_ws), _op(d0), _op(d1), _op(d2),
         _ptr(dfc1), _ptr(dln_g), _ptr(dln_b), _ptr(dfc2), stream_ptr())


def rte_bwd(dvec_sums, emb_table, emb_idx, lin_w, dlin_w, dlin_b, demb):
    n = emb_idx.numel()
    call("a2x_rte_bwd", _ptr(dvec_sums), c_int(n), c_int(lin_w.shape[0]), _ptr(emb_table), _ptr(emb_idx), _ptr(lin_w),
         _ptr(dlin_w), _ptr(dlin_b), _ptr(demb), stream_ptr())


# label generation -------------------------------------------------------------------------------------------------------
def assign_targets(anchor_standup, anchors, gt_standup, gt_boxes, gt_class, gt_offsets, B, pos_threshold, neg_threshold,
                   code_ws, best_ws, targets, pos, neg, class_ids):
    """see include/airv2x_b200.h: a2x_assign_targets"""
    call("a2x_assign_targets", _ptr(anchor_standup), _ptr(anchors), c_int(anchors.shape[0]), _ptr(gt_standup), _ptr(gt_boxes),
         _ptr(gt_class), _ptr(gt_offsets), c_int(0 if gt_boxes is None else gt_boxes.shape[0]), c_int(B),
         c_f(float(pos_threshold)), c_f(float(neg_threshold)), _ptr(code_ws), _ptr(best_ws), _ptr(targets), _ptr(pos),
         _ptr(neg), _ptr(class_ids), stream_ptr())



# Lift-Splat camera branch -------------------------------------------------------------------------------------------------
def lift_splat_fwd(depth, feat, geom, B, N, origin3, dx3, nx3, bev, cells_ws):
    """depth [B*N,D,fH,fW], feat [B*N,C,fH,fW], geom [B,N,D,fH,fW,3] -> bev NHWC [B,ny,nx,nz*C] (see airv2x_b200.h)"""
    _, D, fH, fW = depth.shape
    C = feat.shape[1]
    assert depth.is_contiguous() and feat.is_contiguous() and geom.is_contiguous() and bev.is_contiguous()
    call("a2x_lift_splat_fwd", _ptr(depth), _ptr(feat), _ptr(geom), c_int(B), c_int(N), c_int(D), c_int(fH), c_int(fW), c_int(C),
         (c_f * 3)(*[float(v) for v in origin3]), (c_f * 3)(*[float(v) for v in dx3]), (c_int * 3)(*[int(v) for v in nx3]),
         _ptr(bev), _ptr(cells_ws), stream_ptr())
    return bev


def lift_splat_bwd(depth, feat, cells_ws, dbev, B, N, ddepth, dfeat):
    _, D, fH, fW = depth.shape
    C = feat.shape[1]
    assert dbev.is_contiguous() and ddepth.is_contiguous() and dfeat.is_contiguous()
    call("a2x_lift_splat_bwd", _ptr(depth), _ptr(feat), _ptr(cells_ws), _ptr(dbev), c_int(B), c_int(N), c_int(D), c_int(fH),
         c_int(fW), c_int(C), _ptr(ddepth), _ptr(dfeat), stream_ptr())
