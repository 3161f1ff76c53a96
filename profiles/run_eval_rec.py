"""Short driver for ncu on the rec-side soft-preference evaluation: python profiles/run_eval_rec.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import kgrec_b200 as K
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = K.TransUPModel(False, 100, 50_000, 50_000, 20, False)
qu = torch.arange(0, 4096, device=dev)
cat = m.soft_catalog()
for _ in range(2):
    m.topk_items(qu, k=10, soft_catalog=cat)
torch.cuda.synchronize()
