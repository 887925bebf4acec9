"""Diagnostic (GPU box): per-parameter relative gradient error of the CUDA path vs the oracle autograd."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: F401  (initialises CUDA before the extension loads)
import cases
from oracle import uc2_oracle as O
from uc2_b200 import itm, model
from uc2_b200.batch import to_device
from uc2_b200.utils import set_dropout

task = sys.argv[1] if len(sys.argv) > 1 else "rank"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = cases.config(layers)
fam = O.Family("vlxlmr")
if task == "rank":
    sd = cases.weights(cfg, "retrieval"); m = itm.VLXLMRForImageTextRetrieval(cfg, 2048); b = cases.batch_rank()
else:
    sd = cases.weights(cfg, "pretrain"); m = model.VLXLMRForPretraining(cfg, 2048, 1601)
    b = {"mlm": cases.batch_mlm, "mrfr": cases.batch_mrfr, "mrc-kl": cases.batch_mrc, "mrc": cases.batch_mrc, "itm": cases.batch_itm}[task](seed=int(os.environ.get("SEED", "7")))
m.load_state_dict(cases.with_aliases(sd, "pretrain" if task != "rank" else "retrieval"), strict=False)
m.cuda().train(); set_dropout(m, 0)
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
if task == "rank":
    ref = O.forward_retrieval(sdg, fam, b); ref.mean().backward()
    out = m(to_device(b, "cuda")); out.mean().backward()
    print("loss ref", ref.detach().numpy().ravel(), "got", out.detach().cpu().numpy().ravel())
else:
    ref = O.forward_pretraining(sdg, fam, b, task); lr = O.pretraining_loss(ref, task); lr.backward()
    out = m(to_device(b, "cuda"), task=task)
    if task == "itm":
        itm_l, (p, n) = out; lg = itm_l.mean() + 0.1 * (p.sum() - n.sum()) / (p.size(0) + n.size(0))
    else:
        lg = out.mean()
    lg.backward()
    print("loss ref", lr.item(), "got", lg.item())
params = dict(m.named_parameters())
rows = []
for n, p in params.items():
    gr = sdg[n].grad
    g = p.grad.detach().cpu()
    if gr is None:
        rows.append((n, 0.0, float(g.norm()), -1.0)); continue
    rel = float((g - gr).norm() / (gr.norm() + 1e-30))
    rows.append((n, float(gr.norm()), float(g.norm()), rel))
for n, a, c, r in rows:
    if "layer." in n and ".layer.0." not in n and f".layer.{layers-1}." not in n: continue
    print(f"{n:70s} ref {a:10.4e} got {c:10.4e} relerr {r:8.4f}")
