"""Short driver for ncu on the TUP kernels: python profiles/run_tup.py [soft|gumbel] [train|eval]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import kgrec_b200 as K
mode = sys.argv[1] if len(sys.argv) > 1 else "soft"
what = sys.argv[2] if len(sys.argv) > 2 else "train"
dev = torch.device("cuda:0")
torch.manual_seed(0)
g = torch.Generator().manual_seed(1)
m = K.TransUPModel(False, 100, 50_000, 50_000, 20, mode == "gumbel")
m.grad_mode = "sparse"
if what == "train":
    n = 262144
    u, i, ni = (torch.randint(0, 50_000, (n,), generator=g, dtype=torch.int32).to(dev) for _ in range(3))
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        l, _, _ = m.rank_loss((u, i), (u, ni), target=-1.0, batch_pos=1024)
        l.sum().backward()
else:
    qu = torch.arange(0, 1024, device=dev)
    for _ in range(2):
        m.topk_items(qu, k=10)
torch.cuda.synchronize()
