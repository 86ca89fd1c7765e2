"""GPU probe: train-mode forward error vs the golden reference, split by whether the top-K mask agrees."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import a2x_import
import w2c_common as C
from oracle import w2c_oracle as O
M = a2x_import.pkg("opencood.models.airv2x_where2com")
cfg, gold = C.load_small()
args = cfg["model_args"]
model = M.Airv2xWhere2com(args)
sd = C.golden_state_dict(model, gold)
model.load_state_dict(sd); model.cuda()
dd = C.golden_scene(cfg, gold); ddc = C.to_device(dd, "cuda")
model.train()
random.seed(int(gold["train_K_seed"]))
with torch.no_grad():
    tout = model(ddc)
keep = {}
random.seed(int(gold["train_K_seed"]))
oo, _ = O.where2com_forward({k: v.clone() for k, v in sd.items()}, args, dd, training=True, keep=keep)
m = [t for k_, t in model.engine.bufs.items() if len(k_) == 3 and k_[0] == "mask"][0].cpu()
mm = (m.unsqueeze(1) != keep["mask"])
print("train mask mismatches", int(mm.sum()), "of", m.numel())
for k in ("psm", "rm", "obj"):
    ref = torch.from_numpy(gold["train_" + k])
    e = (tout[k].detach().cpu() - ref).abs()
    print("train %s max %.3e  p99.9 %.3e  median %.3e  n>1e-3: %d  (oracle-vs-golden %.1e)" % (
        k, e.max().item(), np.percentile(e.numpy(), 99.9), e.median().item(), int((e > 1e-3).sum()),
        (oo[k] - ref).abs().max().item()))
smooth = [t for k_, t in model.engine.bufs.items() if len(k_) == 3 and k_[0] == "smooth"][0].cpu()
if "smooth" in keep:
    print("smooth err", (smooth.unsqueeze(1) - keep["smooth"]).abs().max().item())
