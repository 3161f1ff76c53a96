"""Throughput of every model's training step and full-catalog evaluation at BASELINE-config shapes
(CUDA events, warm, one GPU):  python profiles/perf_models.py > profiles/rNN_models_perf.json"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import numpy as np
import torch
import bench
import kgrec_b200 as K
from kgrec_b200 import functional as KF

dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(1)
D, NB, B, KN = 100, 256, 1024, 10


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
ix = [x.to(dev) for x in bench.make_indices(torch, gen, NB)]
pos, neg, corrupt = tuple(ix[:3]), tuple(ix[3:6]), ix[6]
n_tri = pos[0].numel() * (1 + KN)

# ---- KG models: 256 batches x (1024 positives + 10 negatives), |E| = 100k, |R| = 500
for name, mk in (("transe", lambda: K.TransEModel(False, D, 100_000, 500)),
                 ("transh", lambda: K.TransHModel(False, D, 100_000, 500))):
    torch.manual_seed(0)
    m = mk()
    m.grad_mode = "sparse"

    def step():
        m.zero_grad(set_to_none=True)
        m.loss_step_corrupt(pos, corrupt, margin=1.0, batch_pos=B)
    t = timeit(step)
    out[name + "_train_step"] = {"path": "k_group_step (one pass)", "ms": t, "triples_per_s": n_tri / t * 1e3}
    E = 1_000_000
    me = type(m)(False, D, E, 500)
    q = torch.randint(0, E, (4096,), generator=gen).to(dev)
    r = torch.randint(0, 500, (4096,), generator=gen).to(dev)
    t = timeit(lambda: me.topk("tail", q, r, k=10), reps=3, warm=1)
    out[name + "_eval_top10"] = {"path": "k_eval_tiled", "ms": t, "queries": 4096, "catalog": E, "pairs_per_s": 4096 * E / t * 1e3}
    del m, me

torch.manual_seed(0)
m = K.TransRModel(False, D, 100_000, 500)
m.grad_mode = "sparse"
nbr = 32                                               # 32 batches: the d x d matrix gradients are 40 KB per triple
rp = tuple(x[:nbr * B].contiguous() for x in pos)
rn = tuple(x[:nbr * B * KN].contiguous() for x in neg)


rc = corrupt[:nbr * B * KN].contiguous()


def step_r():
    m.zero_grad(set_to_none=True)
    m.loss_step_corrupt(rp, rc, margin=1.0, batch_pos=B)
t = timeit(step_r, reps=3, warm=1)
out["transr_train_step"] = {"path": "k_group_step_r (one pass per group)", "ms": t, "batches": nbr,
                            "triples_per_s": nbr * B * (1 + KN) / t * 1e3}


def step_r_generic():
    m.zero_grad(set_to_none=True)
    l, _, _ = m.rank_loss(rp, rn, margin=1.0, batch_pos=B)
    l.sum().backward()
t = timeit(step_r_generic, reps=3, warm=1)
out["transr_train_step_generic"] = {"path": "k_rank_loss_fwd + k_score_bwd (generic triple format)", "ms": t, "batches": nbr,
                                    "triples_per_s": nbr * B * (1 + KN) / t * 1e3}
q = torch.randint(0, 100_000, (1024,), generator=gen).to(dev)
r = torch.randint(0, 8, (1024,), generator=gen).to(dev)            # 8 distinct relations in the query batch
t = timeit(lambda: m.topk("tail", q, r, k=10), reps=2, warm=1)
out["transr_eval_top10"] = {"path": "library GEMM per distinct relation + k_eval_tiled on explicit query vectors", "ms": t,
                            "queries": 1024, "catalog": 100_000, "distinct_relations": 8, "pairs_per_s": 1024 * 100_000 / t * 1e3}
del m

# ---- rec models: 256 batches x (1024 positives + 1 negative), 50k users x 50k items, P = 20
n_pos = NB * B
u, i, ni = (torch.randint(0, 50_000, (n_pos,), generator=gen, dtype=torch.int32).to(dev) for _ in range(3))
for name in ("tup_soft", "tup_st_gumbel", "ktup_soft", "ktup_st_gumbel"):
    gum = name.endswith("gumbel")
    torch.manual_seed(0)
    if name.startswith("tup"):
        m = K.TransUPModel(False, D, 50_000, 50_000, 20, gum)
    else:
        n_item, n_ent = 50_000, 500_000
        ents = np.random.RandomState(0).permutation(n_ent)[:n_item]
        new_map = {j: (int(ents[j]) if j % 10 < 7 else -1, j) for j in range(n_item)}
        m = K.jTransUPModel(False, D, 50_000, n_item, n_ent, 20, {j: j for j in range(n_item)}, new_map, False, gum)
    m.grad_mode = "sparse"

    def step_rec():
        m.zero_grad(set_to_none=True)
        m.loss_step((u, i), (u, ni), target=-1.0, batch_pos=B)
    t = timeit(step_rec)
    out[name + "_train_step"] = {"path": "k_rec_tile<step> (one pass)", "ms": t, "pairs_per_s": 2 * n_pos / t * 1e3}
    qu = torch.arange(0, 4096, device=dev)
    if gum:
        t = timeit(lambda: m.topk_items(qu[:1024], k=10), reps=2, warm=1)
        out[name + "_eval_top10"] = {"path": "k_eval (pair-specific arg-max preference)", "ms": t, "users": 1024, "items": 50_000,
                                     "pairs_per_s": 1024 * 50_000 / t * 1e3}
    else:
        cat = m.soft_catalog()
        t = timeit(lambda: m.topk_items(qu, k=10, soft_catalog=cat), reps=3, warm=1)
        out[name + "_eval_top10"] = {"path": "k_pref_aug (users) + k_eval_soft", "ms": t, "users": 4096, "items": 50_000,
                                     "pairs_per_s": 4096 * 50_000 / t * 1e3}
    if name.startswith("ktup"):
        kpos = tuple(x % 100_000 if j < 2 else x % 20 for j, x in enumerate(pos))
        kc = torch.where(corrupt < 0, ~((~corrupt) % 100_000), corrupt % 100_000).to(torch.int32)

        def step_kg():
            m.zero_grad(set_to_none=True)
            m.kg_loss_step_corrupt(kpos, kc, margin=1.0, batch_pos=B)
        t = timeit(step_kg)
        out[name + "_kg_train_step"] = {"path": "k_group_step (TransH branch on the joint tables)", "ms": t,
                                        "triples_per_s": n_tri / t * 1e3}
    del m
print(json.dumps(out, indent=1))
