"""Short driver for ncu: one fused forward+backward step at bench size and one eval pass.
    ncu ... python profiles/run_kernels.py [train|eval|all]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import bench
import kgrec_b200 as K

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
torch.manual_seed(0)
gen = torch.Generator().manual_seed(1)
m = K.TransEModel(False, bench.D, bench.N_ENT, bench.N_REL)
m.grad_mode = "sparse"
if what in ("train", "all"):
    ix = [x.to(dev) for x in bench.make_indices(torch, gen, 256)]
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        m.loss_step_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=bench.BATCH)
    torch.cuda.synchronize()
if what in ("eval", "all"):
    q = torch.randint(0, bench.N_ENT, (4096,), generator=gen).to(dev)
    r = torch.randint(0, bench.N_REL, (4096,), generator=gen).to(dev)
    for _ in range(2):
        m.topk("tail", q, r, k=10)
    torch.cuda.synchronize()
