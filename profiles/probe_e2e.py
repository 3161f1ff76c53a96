"""Where does the pipelined end-to-end step spend its time?  (development probe)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import bench
import kgrec_b200 as K
from kgrec_b200.data import DevicePrefetcher

dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(1)
host_sets = [[x.pin_memory() for x in bench.make_indices(torch, gen, 256)] for _ in range(4)]
model = K.TransEModel(False, 100, 100_000, 500)
model.grad_mode = "sparse"
loss_host = torch.empty(256, dtype=torch.float32).pin_memory()
N = 40


def batches(n):
    for s in range(n):
        hs = host_sets[s % 4]
        yield [hs[0], hs[1], hs[2], hs[6]]


def timed(fn, label):
    fn(5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(N)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / N * 1e3
    print(f"{label}: {dt:.3f} ms/step", flush=True)


def copies_only(n):
    for ix in DevicePrefetcher(batches(n), dev):
        pass


def compute_only(n):
    ix = [x.to(dev) for x in next(batches(1))]
    for _ in range(n):
        model.zero_grad(set_to_none=True)
        loss, _, _ = model.loss_step_corrupt(tuple(ix[:3]), ix[3], margin=1.0, batch_pos=1024)
        loss_host.copy_(loss.detach(), non_blocking=True)


def python_only(n):
    ix = [x.to(dev) for x in next(batches(1))]
    t0 = time.perf_counter()
    for _ in range(n):
        model.zero_grad(set_to_none=True)
        loss, _, _ = model.loss_step_corrupt(tuple(ix[:3]), ix[3], margin=1.0, batch_pos=1024)
    dt = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    print(f"   host time to enqueue one step: {dt:.3f} ms", flush=True)


def pipelined(depth):
    def run(n):
        for ix in DevicePrefetcher(batches(n), dev, depth=depth):
            model.zero_grad(set_to_none=True)
            loss, _, _ = model.loss_step_corrupt(tuple(ix[:3]), ix[3], margin=1.0, batch_pos=1024)
            loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return run


timed(copies_only, "H2D copies only (prefetcher, no compute)")
timed(compute_only, "compute only (ids resident)")
python_only(N)
for d in (1, 2, 4):
    timed(pipelined(d), f"pipelined depth={d}")
