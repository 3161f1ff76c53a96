"""TUP / KTUP training-step timing (CUDA events): python profiles/perf_tup.py
Runs the fused ranking-loss forward + backward on n_pos positives x 1 negative (cfg#3 / cfg#4 shapes)
with the tile engine (default) and with the one-warp-per-pair kernels (KGREC_REC_TILE=0)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import numpy as np
import torch
import kgrec_b200 as K

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
n_pos = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
out = {}
for name in ("tup_soft", "tup_gumbel", "ktup_soft", "ktup_gumbel"):
    gum = name.endswith("gumbel")
    torch.manual_seed(0)
    if name.startswith("tup"):
        m = K.TransUPModel(False, 100, 50_000, 50_000, 20, gum)
    else:
        I, E = 50_000, 500_000
        rng = np.random.RandomState(0)
        ents = rng.permutation(E)[:I]
        new_map = {i: (int(ents[i]) if i % 10 < 7 else -1, i) for i in range(I)}
        m = K.jTransUPModel(False, 100, 50_000, I, E, 20, {i: i for i in range(I)}, new_map, False, gum)
    m.grad_mode = "sparse"
    u, i, ni = (torch.randint(0, 50_000, (n_pos,), generator=g, dtype=torch.int32).to(dev) for _ in range(3))
    for eng in ("step", "tile", "warp"):
        if eng == "warp":
            os.environ["KGREC_REC_TILE"] = "0"
        else:
            os.environ.pop("KGREC_REC_TILE", None)
        def step():
            m.zero_grad(set_to_none=True)
            if eng == "step":
                m.loss_step((u, i), (u, ni), target=-1.0, batch_pos=1024)
                return
            l, _, _ = m.rank_loss((u, i), (u, ni), target=-1.0, batch_pos=1024)
            l.sum().backward()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        its = 10
        e0.record()
        for _ in range(its):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / its
        out[f"{name}_{eng}"] = {"ms_per_step": round(ms, 4), "pairs_per_s": round(2 * n_pos / ms * 1e3, 0)}
        print(name, eng, out[f"{name}_{eng}"], flush=True)
print(json.dumps(out))
