"""Small invocations of the kernels added in round 2 for `compute-sanitizer --tool memcheck python profiles/sanitize_targets.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import numpy as np, torch
import kgrec_b200 as K
from kgrec_b200.optim import SparseRowOptimizer
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
lt = lambda a: torch.as_tensor(np.asarray(a), device=dev)
# TransR relation-run step, dense and sparse gradients, two tile configurations
for d, R in ((100, 3), (128, 2), (36, 2)):
    m = K.TransRModel(False, d, 400, R)
    n_pos, Kn = 211, 3
    h, t, r = rng.randint(0, 400, n_pos), rng.randint(0, 400, n_pos), rng.randint(0, R, n_pos)
    ce = rng.randint(0, 400, n_pos * Kn)
    corrupt = lt(np.where(rng.rand(n_pos * Kn) < 0.5, ~ce, ce).astype(np.int32))
    for gm in ("dense", "sparse"):
        m.grad_mode = gm
        m.zero_grad()
        m.loss_step_corrupt((lt(h), lt(t), lt(r)), corrupt, margin=1.0, batch_pos=64)
    opt = SparseRowOptimizer(m, "Adagrad", lr=0.01, clip=5.0)
    opt.step_corrupt((lt(h), lt(t), lt(r)), corrupt, margin=1.0, batch_pos=64)
    q = lt(rng.randint(0, 400, 50)); qr = lt(rng.randint(0, R, 50))
    m.topk("tail", q, qr, k=5)
# row-factored rec steps (soft and ST-Gumbel), TUP and KTUP
os.environ["KGREC_REC_ROWS"] = "force"
for gum in (False, True):
    m = K.TransUPModel(False, 64, 300, 200, 11, gum)
    opt = SparseRowOptimizer(m, "Adagrad", lr=0.01, clip=5.0)
    u, pi, ni = (lt(rng.randint(0, n, 500).astype(np.int32)) for n in (300, 200, 200))
    for _ in range(2):
        opt.step_pairs((u, pi), (u, ni), target=-1.0, batch_pos=100, reg=True)
    cat = m.gumbel_catalog() if gum else m.soft_catalog()
    m.topk_items(lt(np.arange(130) % 300), k=10, soft_catalog=cat)
# tiled top-K with several pieces per query tile (shared bound) and rank counts
m = K.TransEModel(False, 128, 60_000, 7)
q = lt(rng.randint(0, 60_000, 700)); qr = lt(rng.randint(0, 7, 700)); gold = lt(rng.randint(0, 60_000, 700))
m.topk("tail", q, qr, k=10)
m.rank_counts("head", q, qr, gold)
torch.cuda.synchronize()
print("sanitize targets done")
