"""Full-catalog KG evaluation vs query count / catalog size / table construction (one GPU)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import kgrec_b200 as K
from kgrec_b200.models.base import device_init
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(1)


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = []
for d, E, clsname in ((128, 5_000_000, "TransEModel"), (128, 5_000_000, "TransHModel"), (100, 1_000_000, "TransEModel"), (64, 1_000_000, "TransEModel")):
    for init in ("device",):
        if init == "device":
            with device_init(dev):
                m = getattr(K, clsname)(False, d, E, 500)
        else:
            torch.manual_seed(0)
            m = K.TransEModel(False, d, E, 500)
        for nq in (4096, 8192):
            q = torch.randint(0, E, (nq,), generator=gen).to(dev)
            r = torch.randint(0, 500, (nq,), generator=gen).to(dev)
            t = timeit(lambda: m.topk("tail", q, r, k=10))
            gold = torch.randint(0, E, (nq,), generator=gen).to(dev)
            gs = torch.rand(nq, device=dev) * 2
            t2 = timeit(lambda: m.rank_counts("tail", q, r, gold, gold_scores=gs))
            row = {"model": clsname, "d": d, "E": E, "init": init, "nq": nq, "topk_ms": t, "rank_ms": t2, "topk_pairs_per_s": nq * E / t * 1e3,
                   "frac_fp32": nq * E / t * 1e3 / (148 * 128 * 1.965e9 / ((2 if clsname == "TransEModel" else 4) * d))}
            print(json.dumps(row), flush=True)
            out.append(row)
        del m
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_eval_scan.json"), "w"), indent=1)
