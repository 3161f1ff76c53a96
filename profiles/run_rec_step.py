"""Short driver for ncu on the rec-side single-pass step: python profiles/run_rec_step.py [soft|gumbel] [tup|ktup]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import numpy as np
import torch
import kgrec_b200 as K
mode = sys.argv[1] if len(sys.argv) > 1 else "soft"
fam = sys.argv[2] if len(sys.argv) > 2 else "tup"
dev = torch.device("cuda:0")
torch.manual_seed(0)
g = torch.Generator().manual_seed(1)
if fam == "tup":
    m = K.TransUPModel(False, 100, 50_000, 50_000, 20, mode == "gumbel")
else:
    ents = np.random.RandomState(0).permutation(500_000)[:50_000]
    new_map = {j: (int(ents[j]) if j % 10 < 7 else -1, j) for j in range(50_000)}
    m = K.jTransUPModel(False, 100, 50_000, 50_000, 500_000, 20, {j: j for j in range(50_000)}, new_map, False, mode == "gumbel")
m.grad_mode = "sparse"
n = 262144
u, i, ni = (torch.randint(0, 50_000, (n,), generator=g, dtype=torch.int32).to(dev) for _ in range(3))
for _ in range(3):
    m.zero_grad(set_to_none=True)
    m.loss_step((u, i), (u, ni), target=-1.0, batch_pos=1024)
torch.cuda.synchronize()
