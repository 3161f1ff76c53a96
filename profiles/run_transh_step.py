"""Short driver for ncu on the TransH single-pass step kernel: python profiles/run_transh_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import bench
import kgrec_b200 as K
dev = torch.device("cuda:0")
torch.manual_seed(0)
gen = torch.Generator().manual_seed(1)
m = K.TransHModel(False, bench.D, bench.N_ENT, bench.N_REL)
m.grad_mode = "sparse"
ix = [x.to(dev) for x in bench.make_indices(torch, gen, 256)]
for _ in range(2):
    m.zero_grad(set_to_none=True)
    m.loss_step_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=bench.BATCH)
torch.cuda.synchronize()
