"""Multi-GPU plumbing: scenes shard data-parallel (one process per GPU, the reference's DDP —
opencood/tools/train.py:161-163, tools/multi_gpu_utils.py:38-48); the only exchange is the gradient average."""
import torch
import torch.distributed as dist


class GradAverager:
    """Flat-buffer all-reduce of parameter gradients (one collective per step instead of one per tensor)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None

    def __call__(self):
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        sizes = [g.numel() for g in grads]
        if self.flat is None or self.flat.numel() != sum(sizes):
            self.flat = torch.empty(sum(sizes), device=grads[0].device, dtype=grads[0].dtype)
        views = list(self.flat.split(sizes))
        torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
        dist.all_reduce(self.flat)
        self.flat.div_(dist.get_world_size())
        torch._foreach_copy_([g.view(-1) for g in grads], views)
