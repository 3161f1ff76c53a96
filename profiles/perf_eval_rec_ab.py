"""Rec-side full-catalog top-10 (TUP soft / ST-Gumbel, configs[2] and configs[4] shapes) for A/B runs of the eval switches
(KGREC_EVAL_ROTATE, KGREC_EVAL_SHARE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import kgrec_b200 as K
from kgrec_b200.models.base import device_init
dev = torch.device("cuda:0")


def timeit(fn, reps=3, warm=2):
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


print("ROTATE", os.environ.get("KGREC_EVAL_ROTATE", "1"), "SHARE", os.environ.get("KGREC_EVAL_SHARE", "1"))
for d, U, I, nq, gum in ((100, 50_000, 50_000, 4096, False), (100, 50_000, 50_000, 4096, True), (128, 1_000_000, 1_000_000, 16384, False)):
    with device_init(dev):
        torch.cuda.manual_seed(11)
        m = K.TransUPModel(False, d, U, I, 20, gum)
    qu = torch.arange(nq, device=dev) % U
    cat = m.gumbel_catalog() if gum else m.soft_catalog()
    ms = timeit(lambda: m.topk_items(qu, k=10, soft_catalog=cat))
    print(f"d={d} {nq} users x {I} items gumbel={gum}: {ms:.3f} ms  {nq * I / ms / 1e6:.3g}e9 pairs/s")
    del m, cat
