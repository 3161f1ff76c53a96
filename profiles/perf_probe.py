"""Quick timing probe of the individual kernels (CUDA events, warm).  Not the bench: a
development aid whose output is copied into profiles/ per round."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import bench
import kgrec_b200 as K

dev = torch.device("cuda:0")
torch.manual_seed(0)
gen = torch.Generator().manual_seed(1)


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
ix = [x.to(dev) for x in bench.make_indices(torch, gen, 256)]
n_tri = ix[0].numel() * 11
for name, cls in (("transe", K.TransEModel), ("transh", K.TransHModel)):
    m = cls(False, 100, 100_000, 500)
    m.grad_mode = "sparse"
    pos, neg = tuple(ix[:3]), tuple(ix[3:6])
    f = timeit(lambda: m.rank_loss(pos, neg, margin=1.0, batch_pos=1024))

    def fb():
        m.zero_grad(set_to_none=True)
        l, _, _ = m.rank_loss(pos, neg, margin=1.0, batch_pos=1024)
        l.sum().backward()
    fbt = timeit(fb)
    out[name + "_train"] = {"fwd_ms": f, "fwd_bwd_ms": fbt, "fwd_Gtriples_s": n_tri / f / 1e6, "bwd_ms": fbt - f}
    for (d, E) in ((100, 100_000), (128, 100_000)):
        for l1 in (False, True):
            me = cls(l1, d, E, 500)
            q = torch.randint(0, E, (4096,), generator=gen).to(dev)
            r = torch.randint(0, 500, (4096,), generator=gen).to(dev)
            t = timeit(lambda: me.topk("tail", q, r, k=10), reps=3, warm=1)
            out["%s_eval_topk_d%d_%s" % (name, d, "l1" if l1 else "l2")] = {"ms": t, "Gpairs_s": 4096 * E / t / 1e6}
    me = cls(False, 100, 100_000, 500)
    t = timeit(lambda: me.evaluateTail(q[:1024], r[:1024]), reps=3, warm=1)
    out[name + "_eval_full_1024q"] = {"ms": t, "Gpairs_s": 1024 * 100_000 / t / 1e6}
for gumbel in (True, False):
    mu = K.TransUPModel(False, 100, 50_000, 50_000, 20, gumbel)
    mu.grad_mode = "sparse"
    u = torch.randint(0, 50_000, (262144,), generator=gen, dtype=torch.int32).to(dev)
    i = torch.randint(0, 50_000, (262144,), generator=gen, dtype=torch.int32).to(dev)
    ni = torch.randint(0, 50_000, (262144,), generator=gen, dtype=torch.int32).to(dev)
    f = timeit(lambda: mu.rank_loss((u, i), (u, ni), target=-1.0, batch_pos=1024))

    def fbu():
        mu.zero_grad(set_to_none=True)
        l, _, _ = mu.rank_loss((u, i), (u, ni), target=-1.0, batch_pos=1024)
        l.sum().backward()
    fbt = timeit(fbu)
    key = "tup_%s" % ("gumbel" if gumbel else "soft")
    out[key + "_train"] = {"fwd_ms": f, "fwd_bwd_ms": fbt, "fwd_Gpairs_s": 2 * 262144 / f / 1e6}
    qu = torch.arange(0, 1024, device=dev)
    t = timeit(lambda: mu.topk_items(qu, k=10), reps=2, warm=1)
    out[key + "_eval_topk"] = {"ms": t, "Gpairs_s": 1024 * 50_000 / t / 1e6}
print(json.dumps(out, indent=1))
