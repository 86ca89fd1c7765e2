"""The caller side of the training step: what `opencood/tools/train.py:191-260` and
`opencood/tools/train_utils.py:35-118, :371-465` do around `model(batch)`, arranged for the B200 path.

  * labels are assigned on the GPU from the padded ground-truth boxes (`labels.TargetAssigner`) instead of arriving as
    fp64 maps from a DataLoader worker;
  * forward + loss + backward is the model's fused `train_step` (or the CUDA-graph replay when the batch carries raw
    point clouds), the optimizer / scheduler are the reference's (`setup_optimizer`, `setup_lr_schedular`: same yaml keys);
  * checkpoints keep the reference's file name and dict layout (`net_epoch%d.pth` with `epoch`, `model_state_dict`,
    `optimizer_state_dict`, `scheduler_state_dict`, train.py:243-253) and RESUME WORKS: the reference's
    `load_saved_model` reads that dict as if it were a flat state_dict (every key is dropped, nothing is loaded,
    train_utils.py:88-116) and its `findLastCheckpoint` returns an undefined name when checkpoints exist (:54-63).

Host logic only; every number-crunching step is a kernel behind the model / assigner. No CPU fallback.
"""
import glob
import os
import re

import torch

from .dist import GradAverager
from .labels import TargetAssigner


def setup_optimizer(hypes, model):
    """train_utils.py:371-390: `getattr(torch.optim, core_method)(model.parameters(), lr=..., **args)`; on CUDA the
    fused (single multi-tensor kernel) implementation of the same optimizer is selected when torch offers it."""
    cfg = hypes["optimizer"]
    cls = getattr(torch.optim, cfg["core_method"], None)
    if cls is None:
        raise ValueError("{} is not supported".format(cfg["core_method"]))
    params = [p for p in model.parameters() if p.requires_grad]
    kw = dict(cfg.get("args", {}))
    if cfg["core_method"] in ("Adam", "AdamW", "SGD") and all(p.is_cuda for p in params):
        kw.setdefault("fused", True)
    return cls(params, lr=cfg["lr"], **kw)


class CosineWarmupLR:
    """`timm.scheduler.cosine_lr.CosineLRScheduler(optimizer, t_initial, lr_min, warmup_lr_init, warmup_t, cycle_limit=1,
    t_in_epochs=False)` as `setup_lr_schedular` builds it for `cosineannealwarm` (train_utils.py:430-447; the shipped
    V2X-R yamls). timm is not a dependency here: the schedule is restated from its published implementation (linear warm-up
    from `warmup_lr_init`, then lr_min + (lr - lr_min) (1 + cos(pi t / t_initial)) / 2, lr_min after one cycle) — parity with
    timm itself is unpinned. Like timm's object, constructing it sets the learning rate to `warmup_lr_init`, and because the
    schedule counts UPDATES (`t_in_epochs=False`) the epoch-level `step(epoch)` that `tools/train.py:289` issues changes
    nothing: only `step_update(num_updates)` moves the rate (the reference's loop never calls it, so the shipped yaml trains
    at the constant warm-up rate; `Trainer(per_iteration_schedule=True)` calls it once per step)."""

    def __init__(self, optimizer, t_initial, lr_min, warmup_lr_init, warmup_t):
        import math
        self._math = math
        self.optimizer = optimizer
        self.t_initial, self.lr_min, self.warmup_lr_init, self.warmup_t = int(t_initial), lr_min, warmup_lr_init, int(warmup_t)
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", g["lr"])
        self.base_values = [g["initial_lr"] for g in optimizer.param_groups]
        self.num_updates = 0
        if self.warmup_t:
            self._set([self.warmup_lr_init for _ in self.base_values])

    def _set(self, values):
        for g, v in zip(self.optimizer.param_groups, values):
            g["lr"] = v

    def lr_at(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * (v - self.warmup_lr_init) / self.warmup_t for v in self.base_values]
        if t // self.t_initial >= 1:                         # cycle_limit = 1
            return [self.lr_min for _ in self.base_values]
        c = 0.5 * (1 + self._math.cos(self._math.pi * (t % self.t_initial) / self.t_initial))
        return [self.lr_min + (v - self.lr_min) * c for v in self.base_values]

    def step(self, epoch=None, metric=None):
        """epoch-level call: a schedule in updates ignores it (timm `_get_values(epoch, on_epoch=True)` -> None)"""

    def step_update(self, num_updates, metric=None):
        self.num_updates = int(num_updates)
        self._set(self.lr_at(self.num_updates))

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k not in ("optimizer", "_math")}

    def load_state_dict(self, state):
        self.__dict__.update(state)


def setup_lr_scheduler(hypes, optimizer, init_epoch=None, n_iter_per_epoch=None):
    """train_utils.py:393-452: step / multistep / exponential on torch's schedulers, cosineannealwarm on `CosineWarmupLR`"""
    cfg = hypes["lr_scheduler"]
    m = cfg["core_method"]
    if m == "cosineannealwarm":
        if not n_iter_per_epoch:
            raise ValueError("lr_scheduler cosineannealwarm counts updates: pass n_iter_per_epoch (len(train_loader), train.py:177-184)")
        return CosineWarmupLR(optimizer, cfg["epoches"] * n_iter_per_epoch, cfg["lr_min"], cfg["warmup_lr"],
                              cfg["warmup_epoches"] * n_iter_per_epoch)
    if m == "step":
        sch = torch.optim.lr_scheduler.StepLR(optimizer, step_size=cfg["step_size"], gamma=cfg["gamma"])
    elif m == "multistep":
        sch = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=cfg["step_size"], gamma=cfg["gamma"])
    elif m == "exponential":
        sch = torch.optim.lr_scheduler.ExponentialLR(optimizer, cfg["gamma"])
    else:
        raise NotImplementedError("lr scheduler %r is not implemented" % m)
    for _ in range(init_epoch or 0):
        sch.step()
    return sch


def find_last_checkpoint(save_dir):
    """highest N among `*epochN.pth` files, 0 if there is none (what train_utils.py:54-63 means to do)"""
    epochs = []
    for f in glob.glob(os.path.join(save_dir, "*epoch*.pth")):
        m = re.findall(r".*epoch(\d+)\.pth$", f)
        if m:
            epochs.append(int(m[0]))
    return max(epochs) if epochs else 0


def load_saved_model(saved_path, model, epoch=None, optimizer=None, scheduler=None):
    """Same call as train_utils.load_saved_model (`initial_epoch, model = load_saved_model(path, model)`), accepting
    both checkpoint layouts: the dict train.py writes and a flat state_dict (the published checkpoints). `module.`
    prefixes of DataParallel / DDP are stripped and shape mismatches skipped like the reference (:88-116)."""
    assert os.path.exists(saved_path), "{} not found".format(saved_path)
    initial_epoch = find_last_checkpoint(saved_path) if epoch is None else int(epoch)
    if initial_epoch == 0:
        return 0, model
    ckpt = torch.load(os.path.join(saved_path, "net_epoch%d.pth" % initial_epoch), map_location="cpu")
    flat = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt
    own = model.state_dict()
    state = {}
    for k, v in flat.items():
        if k.startswith("module") and not k.startswith("module_list"):
            k = k[7:]
        if k in own and tuple(own[k].shape) == tuple(v.shape):
            state[k] = v
    model.load_state_dict(state, strict=False)
    if isinstance(ckpt, dict) and "model_state_dict" in ckpt:
        if optimizer is not None and "optimizer_state_dict" in ckpt:
            optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if scheduler is not None and "scheduler_state_dict" in ckpt:
            scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    return initial_epoch, model


class Trainer:
    """`Trainer(model, hypes)`; `loss3 = trainer.step(batch)` with batch = the model's data_dict plus the padded boxes
    `object_bbx_center [B,max_num,7]`, `object_bbx_mask [B,max_num]`, `object_class_ids [B,max_num]` (the tensors the
    dataset hands to `generate_label_airv2x`) — or a ready `label_dict`. Returns the device tensor [reg, cls, obj]."""

    def __init__(self, model, hypes, graph=True, n_iter_per_epoch=None, per_iteration_schedule=False):
        self.model, self.hypes = model, hypes
        self.iteration, self.per_iteration_schedule = 0, per_iteration_schedule
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("Trainer (B200) needs the model on a CUDA device; there is no CPU path")
        self.assigner = TargetAssigner(hypes["postprocess"], dev)
        self.optimizer = setup_optimizer(hypes, model)
        self.scheduler = setup_lr_scheduler(hypes, self.optimizer, n_iter_per_epoch=n_iter_per_epoch)
        la = hypes["loss"]["args"]
        self.cls_weight, self.reg_coe = float(la["cls_weight"]), float(la["reg"])
        self.graph = graph
        self.epoch = 0
        # scene-parallel training (one process per GPU, tools/train.py:161-163 wraps the model in DDP): one flat all-reduce
        # of the gradients per step; a no-op unless torch.distributed is initialised with world_size > 1
        # (GradAverager: p.grad become views of one flat buffer); models with attach_grad_sync() reduce INSIDE their fused
        # step, overlapped with the tail of the backward
        if hasattr(model, "attach_grad_sync") and type(model).__name__ == "Airv2xWhere2com":
            model.attach_grad_sync()
            self.average_grads = lambda: None
        else:
            self.average_grads = GradAverager(model.parameters())

    def labels(self, batch):
        if "label_dict" in batch and "targets" in batch["label_dict"]:      # ready label maps (the reference's dataset)
            return batch["label_dict"]
        return self.assigner(batch["object_bbx_center"], batch["object_bbx_mask"], batch["object_class_ids"])

    def step(self, batch, **kw):
        model = self.model
        model.train()
        labels = self.labels(batch)
        data = {k: v for k, v in batch.items() if k not in ("label_dict", "object_bbx_center", "object_bbx_mask", "object_class_ids")}
        graphed = self.graph and data.get("raw_points") is not None \
            and hasattr(model, "train_step_graphed") and type(model).__name__ == "Airv2xWhere2com"
        if graphed:
            loss3 = model.train_step_graphed(data, labels, self.cls_weight, self.reg_coe)
        else:
            loss3 = model.train_step(data, labels, self.cls_weight, self.reg_coe, **kw)
        self.average_grads()                        # gradients were written into p.grad by the fused step
        if self.per_iteration_schedule and hasattr(self.scheduler, "step_update"):
            self.scheduler.step_update(self.iteration)
        self.optimizer.step()
        self.iteration += 1
        return loss3

    def end_epoch(self, saved_path=None):
        """scheduler step + checkpoint. Scene-parallel runs: the BatchNorm running statistics (the only buffers; every
        rank sees different scenes) are broadcast from rank 0 first, like DDP's buffer broadcast (tools/train.py:161-163),
        rank 0 alone writes the file and every rank waits for it."""
        self.scheduler.step()
        self.epoch += 1
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if multi:
            for b in self.model.buffers():
                if b.is_floating_point():
                    dist.broadcast(b, 0)
        if saved_path is not None:
            if not multi or dist.get_rank() == 0:
                os.makedirs(saved_path, exist_ok=True)
                torch.save({"epoch": self.epoch - 1, "model_state_dict": self.model.state_dict(),
                            "optimizer_state_dict": self.optimizer.state_dict(),
                            "scheduler_state_dict": self.scheduler.state_dict()},
                           os.path.join(saved_path, "net_epoch%d.pth" % self.epoch))
            if multi:
                dist.barrier()

    def resume(self, saved_path):
        self.epoch, _ = load_saved_model(saved_path, self.model, None, self.optimizer, self.scheduler)
        if hasattr(self.scheduler, "num_updates") and self.scheduler.num_updates:
            self.iteration = self.scheduler.num_updates + 1          # an update-counting schedule continues where it stopped
        return self.epoch


class Evaluator:
    """The evaluation loop of `tools/inference_multi_scenario.py:330-432` / `tools/inference_utils.py:99-134` for the
    intermediate-fusion models: eval forward -> `dataset.post_process` (GPU decode + rotated NMS, ground truth of the ego) ->
    TP / FP bookkeeping at IoU 0.3 / 0.5 / 0.7 (`caluclate_tp_fp`, utils/eval_utils_opv2v.py:41-95) -> VOC AP
    (`eval_final_results` :155-189; like the reference's call, detections stay in frame order unless `global_sort`).
    `scenario_of(batch)` groups the statistics like the reference's per-scenario timestamp key (default: one group)."""

    def __init__(self, model, dataset, ious=(0.3, 0.5, 0.7), scenario_of=None):
        self.model, self.dataset, self.ious = model, dataset, tuple(ious)
        self.scenario_of = scenario_of or (lambda batch: "all")
        self.stats, self.comm_rates = {}, []

    def step(self, batch):
        """batch: the collated dict of `collate_batch_test` ({"ego": ...}); returns this sample's predictions"""
        from .postprocess import calculate_tp_fp
        self.model.eval()
        with torch.no_grad():
            out = self.model(batch["ego"])
        pred_box, score, labels, boxes3d, gt_box, gt_cls, gt_ids = self.dataset.post_process(batch, {"ego": out})
        if "comm_rate" in out:
            self.comm_rates.append(float(out["comm_rate"]))
        st = self.stats.setdefault(self.scenario_of(batch), {t: {"tp": [], "fp": [], "gt": 0, "score": []} for t in self.ious})
        if pred_box is not None and pred_box.shape[0] > 0:   # the reference skips a frame without detections (:366-367)
            for t in self.ious:
                calculate_tp_fp(pred_box, score, gt_box, st, t)
        return pred_box, score, labels, boxes3d, gt_box

    def summary(self, global_sort=False):
        """{scenario: {iou: AP}} plus the mean communication rate"""
        from .postprocess import calculate_ap
        out = {}
        for name, st in self.stats.items():
            out[name] = {t: (calculate_ap(st, t, global_sort)[0] if st[t]["gt"] > 0 and st[t]["tp"] else 0.0) for t in self.ious}
        out["comm_rate"] = sum(self.comm_rates) / len(self.comm_rates) if self.comm_rates else 0
        return out
