"""The fused forward / backward of the drop-in models as `torch.library` custom ops, so that autograd,
autocast, DDP and torch.compile see ORDINARY dispatcher ops instead of an opaque autograd.Function:

    a2x::fused_forward(int key, Tensor[] params, int[] out_shape) -> Tensor      (CUDA impl + Meta/fake impl + autograd)
    a2x::fused_backward(int key, Tensor dheads, Tensor[] params) -> Tensor[]     (CUDA impl + Meta/fake impl)

`params` are the trainable parameters (the differentiable inputs), the output is the NHWC head-logit tensor; the scene
(point clouds, layout, per-step state) travels through a Python-side context looked up by `key` (one per model
instance), because it is not tensor-shaped data the dispatcher could carry. The reference's training loop
(opencood/tools/train.py:216-221: `model(batch)` -> criterion -> `loss.backward()`), its DDP wrapper (:161-163) and
`torch.autocast` regions work unchanged: the ops are registered for CUDA only (there is no CPU implementation — a CPU
tensor raises from the dispatcher), take and return fp32, and under torch.compile they trace as two graph nodes through
their fake implementations.
"""
import torch

_CTX = {}


def bind(model, run_forward, run_backward):
    """(re)bind the per-call closures of `model`; returns the integer key the ops take"""
    key = id(model)
    _CTX[key] = (run_forward, run_backward)
    return key


@torch.library.custom_op("a2x::fused_forward", mutates_args=(), device_types="cuda")
def fused_forward(key: int, params: list[torch.Tensor], out_shape: list[int]) -> torch.Tensor:
    heads = _CTX[key][0]()
    assert list(heads.shape) == list(out_shape), (heads.shape, out_shape)
    return heads.clone()      # the engine owns `heads`; autograd / the caller get their own tensor


@fused_forward.register_fake
def _(key, params, out_shape):
    return params[0].new_empty(out_shape)


@torch.library.custom_op("a2x::fused_backward", mutates_args=(), device_types="cuda")
def fused_backward(key: int, dheads: torch.Tensor, params: list[torch.Tensor]) -> list[torch.Tensor]:
    return _CTX[key][1](dheads.contiguous())


@fused_backward.register_fake
def _(key, dheads, params):
    return [torch.empty_like(p) for p in params]


def _setup(ctx, inputs, output):
    ctx.key = inputs[0]
    ctx.save_for_backward(*inputs[1])


def _backward(ctx, dheads):
    grads = torch.ops.a2x.fused_backward(ctx.key, dheads, list(ctx.saved_tensors))
    return None, grads, None


fused_forward.register_autograd(_backward, setup_context=_setup)
