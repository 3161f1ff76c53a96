"""BASELINE configs[4] shapes on one GPU: d=128 full-catalog evaluation, 5M entities (KG side) and 1M items
(rec side, TUP soft preferences), top-10 for a 4096-query slice (time per slice; the full 1M-user pass is
244 such slices):  python profiles/perf_cfg5.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import kgrec_b200 as K

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


out = {}
d, nq = 128, 4096
for name, cls in (("transe", K.TransEModel), ("transh", K.TransHModel)):
    E = 5_000_000
    torch.manual_seed(0)
    m = cls(False, d, E, 500)
    q = torch.randint(0, E, (nq,), generator=gen).to(dev)
    r = torch.randint(0, 500, (nq,), generator=gen).to(dev)
    t = timeit(lambda: m.topk("tail", q, r, k=10))
    out[name + "_kg_top10_d128_5M"] = {"ms": t, "queries": nq, "catalog": E, "pairs_per_s": nq * E / t * 1e3,
                                       "frac_of_fp32_bound": nq * E / t * 1e3 / (148 * 128 * 1.965e9 / ((2 if name == "transe" else 4) * d))}
    del m
torch.manual_seed(0)
I = 1_000_000
m = K.TransUPModel(False, d, 1_000_000, I, 20, False)
qu = torch.randint(0, 1_000_000, (nq,), generator=gen).to(dev)
cat = m.soft_catalog()
t = timeit(lambda: m.topk_items(qu, k=10, soft_catalog=cat))
out["tup_soft_rec_top10_d128_1M"] = {"ms": t, "users": nq, "items": I, "pairs_per_s": nq * I / t * 1e3,
                                     "frac_of_fp32_bound": nq * I / t * 1e3 / (148 * 128 * 1.965e9 / (6 * d))}
print(json.dumps(out, indent=1))
