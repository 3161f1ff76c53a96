import os, sys, torch
sys.path[:0] = ["/root/repo", "/root/repo/joint-kg-recommender_b200"]
import kgrec_b200 as K
from kgrec_b200.models.base import device_init
from kgrec_b200.optim import SparseRowOptimizer
dev = torch.device("cuda:0")
def timeit(fn, reps=3, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
torch.manual_seed(13)
with device_init(dev):
    tm = K.TransUPModel(False, 100, 50_000, 50_000, 20, True)
n_pos = 256 * 1024
tg = torch.Generator().manual_seed(5)
tu, ti, tn = (torch.randint(0, 50_000, (n_pos,), generator=tg, dtype=torch.int32).to(dev) for _ in range(3))
qu = torch.arange(4096, device=dev) % 50_000
def ev(tag):
    gcat = tm.gumbel_catalog()
    g = timeit(lambda: tm.topk_items(qu, k=10, soft_catalog=gcat))
    tm.use_st_gumbel = False
    sc = tm.soft_catalog()
    s = timeit(lambda: tm.topk_items(qu, k=10, soft_catalog=sc))
    tm.use_st_gumbel = True
    st = {k: (float(v.abs().max()), bool(torch.isfinite(v).all())) for k, v in tm._weights().items()}
    print(tag, "gumbel eval ms %.2f soft eval ms %.2f" % (g, s), st)
ev("fresh")
gopt = SparseRowOptimizer(tm, optimizer_type="Adagrad", lr=0.005, clip=5.0)
for i in range(7):
    gopt.step_pairs((tu, ti), (tu, tn), target=-1.0, batch_pos=1024, reg=True)
ev("after 7 rows-path steps")
os.environ["KGREC_REC_ROWS"] = "0"
for i in range(7):
    gopt.step_pairs((tu, ti), (tu, tn), target=-1.0, batch_pos=1024, reg=True)
ev("after 7 more pair-kernel steps")
