"""k_group_step A/B timing (CUDA events): KGREC_GROUP_STEP = 0 (general kernel) | n (issue-optimised kernel, no row prefetch) | other (default)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import bench
import kgrec_b200 as K

dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(1)
sets = [[x.to(dev) for x in bench.make_indices(torch, gen, 256)] for _ in range(3)]
n_tri = sets[0][0].numel() * 11
for l1 in (False, True):
    torch.manual_seed(0)
    m = K.TransEModel(l1, 100, 100_000, 500)
    m.grad_mode = "sparse"
    for env in ("n", "4"):
        os.environ["KGREC_GROUP_STEP"] = env
        def step(s):
            ix = sets[s % 3]
            m.zero_grad(set_to_none=True)
            m.loss_step_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=1024)
        for s in range(3):
            step(s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(20):
            step(s)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(f"l1={l1} KGREC_GROUP_STEP={env}: {ms:.4f} ms/step  {n_tri / ms * 1e3 / 1e9:.3f} G triples/s", flush=True)
for l1 in (False, True):
    torch.manual_seed(0)
    m = K.TransHModel(l1, 100, 100_000, 500)
    m.grad_mode = "sparse"
    for env in ("0", "n", "3"):
        os.environ["KGREC_GROUP_STEP"] = env
        def step(s):
            ix = sets[s % 3]
            m.zero_grad(set_to_none=True)
            m.loss_step_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=1024)
        for s in range(3):
            step(s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(20):
            step(s)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(f"transh l1={l1} KGREC_GROUP_STEP={env}: {ms:.4f} ms/step  {n_tri / ms * 1e3 / 1e9:.3f} G triples/s", flush=True)
# the autograd pair: kgrec_corrupt_loss_fwd + kgrec_corrupt_loss_bwd (the step kernel in backward mode)
for cls in (K.TransEModel, K.TransHModel):
    torch.manual_seed(0)
    m = cls(False, 100, 100_000, 500)
    m.grad_mode = "sparse"
    for env in ("0", "4"):
        os.environ["KGREC_GROUP_STEP"] = env
        def step(s):
            ix = sets[s % 3]
            m.zero_grad(set_to_none=True)
            l, _, _ = m.rank_loss_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=1024)
            l.sum().backward()
        for s in range(3):
            step(s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(20):
            step(s)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(f"autograd fwd+bwd {cls.__name__} KGREC_GROUP_STEP={env}: {ms:.4f} ms/step  {n_tri / ms * 1e3 / 1e9:.3f} G triples/s", flush=True)
