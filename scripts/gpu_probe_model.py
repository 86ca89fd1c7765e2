"""GPU probe: full Where2comm model (eval + train fwd/bwd) vs golden fixtures and the oracle. Report only."""
import os, sys, random, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import a2x_import
import w2c_common as C
from oracle import w2c_oracle as O
M = a2x_import.pkg("opencood.models.airv2x_where2com")
prec = sys.argv[1] if len(sys.argv) > 1 else "split3"
cfg, gold = C.load_small()
args = cfg["model_args"]
model = M.Airv2xWhere2com(args, precision=prec)
sd = C.golden_state_dict(model, gold)
model.load_state_dict(sd)
model.cuda()
dd = C.golden_scene(cfg, gold)
ddc = C.to_device(dd, "cuda")
model.eval()
with torch.no_grad():
    out = model(ddc)
torch.cuda.synchronize()
for k in ("psm", "rm", "obj"):
    ref = torch.from_numpy(gold["eval_" + k])
    err = (out[k].cpu() - ref).abs().max().item()
    print("[%s] eval %s max abs err vs reference golden: %.3e (max|ref| %.2f)" % (prec, k, err, ref.abs().max().item()))
print("eval com", float(out["com"]), "golden", float(gold["eval_com"]), "comm_rate", out["comm_rate"], int(gold["eval_comm_rate"]))

# stage-level check vs the oracle (CPU)
keep = {}
with torch.no_grad():
    oo, _ = O.where2com_forward(sd, args, dd, training=False, keep=keep)
eng = model.engine
def nchw(t): return t.permute(0, 3, 1, 2).cpu()
def fullbuf(name, shape):
    for (n, s, dt), t in eng.bufs.items():
        if n == name: return t
    return None
canv = fullbuf("canvas", None)
cv = canv
print("canvas err", (nchw(cv) - keep["spatial_features"]).abs().max().item(), "max", keep["spatial_features"].abs().max().item())
m = fullbuf("mask", None)
print("mask mismatches", int((m.cpu().unsqueeze(1) != keep["mask"]).sum()), "of", m.numel())
for i in range(3):
    f = fullbuf("B.fuse%d" % i, None)
    fv = f
    print("fused level", i, "err", (nchw(fv) - keep["fused_l%d" % i]).abs().max().item(), "max", keep["fused_l%d" % i].abs().max().item())

# ---- train step through autograd + oracle loss
try:
    model.train(); model.load_state_dict(sd)
    H, W = out["psm"].shape[2:]
    labels = O.make_labels(int(gold["label_seed"]), 1, H, W, args["anchor_number"])
    random.seed(int(gold["train_K_seed"]))
    tout = model(ddc)
    for k in ("psm", "rm", "obj"):
        ref = torch.from_numpy(gold["train_" + k])
        print("[%s] train %s max abs err: %.3e" % (prec, k, (tout[k].detach().cpu() - ref).abs().max().item()))
    print("train com", float(tout["com"]), "golden", float(gold["train_com"]))
    lab_c = {k: v.cuda() for k, v in labels.items()}
    loss, lr, lc, lo = O.point_pillar_loss_multiclass(tout, {k: v for k, v in lab_c.items()}, args["num_class"],
                                                      cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"]) \
        if False else (None, None, None, None)
except Exception:
    traceback.print_exc()

# oracle loss runs on CPU tensors; move outputs
try:
    cpu_out = {k: tout[k].cpu() for k in ("psm", "rm", "obj")}
    # autograd through .cpu() is fine
    loss, lr, lc, lo = O.point_pillar_loss_multiclass(cpu_out, labels, args["num_class"], cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])
    print("train loss %.8f golden %.8f" % (float(loss), float(gold["train_loss"])))
    model.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    worst = []
    for n, p in model.named_parameters():
        key = "grad_" + n
        if key not in gold.files: continue
        g = p.grad
        if g is None:
            print("NO GRAD", n); continue
        ref = gold[key]; got = C.sample(g, 512)
        denom = max(np.abs(ref).max(), 1e-12)
        worst.append((np.abs(got - ref).max() / denom, n, float(np.abs(ref).max())))
    worst.sort(reverse=True)
    for w in (worst if os.environ.get("ALLG") else worst[:12]): print("grad rel err %.3e  %s (max|ref| %.3e)" % w)
    print("median grad rel err %.3e over %d params" % (np.median([w[0] for w in worst]), len(worst)))
    # fused loss kernel vs oracle loss
    heads = eng.saved["heads"]
    lab_k = {"targets": lab_c["targets"].float().contiguous(), "pos_equal_one": lab_c["pos_equal_one"].float().contiguous(),
             "class_ids": lab_c["class_ids"].int().contiguous()}
    loss3, dheads = eng.loss(heads, lab_k, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])
    torch.cuda.synchronize()
    print("fused loss (reg, cls, obj)", loss3.tolist(), "oracle", float(lr), float(lc), float(lo))
    # running stats
    wb = 0.0
    for n, b in model.named_buffers():
        key = "buf_" + n
        if key in gold.files:
            wb = max(wb, float(np.abs(C.sample(b, 64) - gold[key]).max()))
    print("running-stat max abs err vs reference after one train step: %.3e" % wb)
except Exception:
    traceback.print_exc()
