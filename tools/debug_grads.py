import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200")); sys.path.insert(1, ROOT); sys.path.insert(2, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from conftest import load_golden, rel_l2
from oracle.weights import synth_state_dict
from models.U_Net import U_Net
name = sys.argv[1] if len(sys.argv) > 1 else "gpu_small"
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
fx = load_golden(f"unet_{name}.pt")
net = U_Net(**fx["kwargs"]); net.load_state_dict(synth_state_dict(fx["shapes"], fx["seed"])); net = net.cuda().train().set_precision(prec)
cond = fx["cond"].cuda() if fx["cond"] is not None else None
out = net(fx["x"].cuda(), fx["t"].cuda(), cond)
loss = F.mse_loss(out, fx["target"].cuda()); loss.backward()
print("loss", float(loss), fx["loss"], "out rel", rel_l2(out.detach().cpu(), fx["out"]))
named = dict(net.named_parameters())
lay = net.engine().layout
order = {id(p): i for i, p in enumerate(lay.params)}
rows = []
for pname, g in fx["grads"].items():
    p = named[pname]
    if p.grad is None: print("NO GRAD", pname); continue
    got = p.grad.detach().float().cpu().flatten()
    en = abs(float(got.norm()) - g["norm"]) / (g["norm"] + 1e-12)
    ef = rel_l2(got, g["full"].flatten()) if g["full"] is not None else rel_l2(got[:64], g["head"])
    rows.append((order.get(id(p), -1), pname, en, ef, g["norm"]))
rows.sort()
for r in rows: print(f"{r[0]:4d} {r[1]:70s} norm_err {r[2]:.2e}  vec_err {r[3]:.2e}  ref_norm {r[4]:.2e}")
