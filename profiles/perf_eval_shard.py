"""What one GPU of an N-GPU sharded evaluation pass does, measured on ONE GPU: the same queries against a 1/N shard of
the catalog.  t(shard) vs t(full) / N separates the part of a pass that does not scale with the shard (launch + merge +
python overhead; the NCCL collective is not in this number) from the kernel's own efficiency on a smaller shard.
    python profiles/perf_eval_shard.py > gpurun_out/eval_shard.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import kgrec_b200 as K
from kgrec_b200.models.base import device_init
dev = torch.device("cuda:0")
d = 128


def timeit(fn, reps=5, warm=2):
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
for n_cat, nq in ((1_000_000, 4096), (5_000_000, 8192)):
    with device_init(dev):
        torch.cuda.manual_seed(7)
        m = K.TransEModel(False, d, n_cat, 500)
    g = torch.Generator().manual_seed(99)
    qh = torch.randint(0, n_cat, (nq,), generator=g).to(dev)
    qr = torch.randint(0, 500, (nq,), generator=g).to(dev)
    gold = torch.randint(0, n_cat, (nq,), generator=g).to(dev)
    gs = m.gold_scores("tail", qh, qr, gold)
    W = m.ent_embeddings.weight.detach()
    res = {}
    for world in (1, 2, 4, 8):
        per = (n_cat + world - 1) // world
        shard = W[:per]
        t_top = timeit(lambda: m.topk("tail", qh, qr, k=10, catalog=shard, id_base=0))
        t_cnt = timeit(lambda: m.rank_counts("tail", qh, qr, gold, gold_scores=gs, catalog=shard, id_base=0))
        res[str(world)] = {"shard_rows": per, "top10_ms": t_top, "rank_counts_ms": t_cnt}
    for world in ("2", "4", "8"):
        for k in ("top10_ms", "rank_counts_ms"):
            res[world][k.replace("_ms", "_eff")] = res["1"][k] / int(world) / res[world][k]
    out["%dq_x_%d" % (nq, n_cat)] = res
    del m, W
print(json.dumps(out, indent=1))
