"""TransR training step: the group kernel (loss_step_corrupt) vs the generic kernels (rank_loss + backward)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import bench
import kgrec_b200 as K
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(1)
nb = 32
ix = [x.to(dev) for x in bench.make_indices(torch, gen, nb)]
pos, neg, corrupt = tuple(ix[:3]), tuple(ix[3:6]), ix[6]
n_tri = pos[0].numel() * 11
m = K.TransRModel(False, 100, 100_000, 500)
m.grad_mode = "sparse"
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def step():
    m.zero_grad(set_to_none=True)
    m.loss_step_corrupt(pos, corrupt, margin=1.0, batch_pos=1024)
def generic():
    m.zero_grad(set_to_none=True)
    l, _, _ = m.rank_loss(pos, neg, margin=1.0, batch_pos=1024)
    l.sum().backward()
a, b = t(step), t(generic)
print(f"transr group step: {a:.3f} ms ({n_tri / a * 1e3:.3g} triples/s)   generic fwd+bwd: {b:.3f} ms ({n_tri / b * 1e3:.3g} triples/s)")
