"""A/B of the TransE step kernel's gather: register loads + L1 prefetch (default) vs rows staged through shared
memory by per-row TMA bulk copies (KGREC_GROUP_STEP=t<w><s>), at tables that are L2-resident (100k entities)
and far larger than L2 (500k, 5M).  The switch is read once per process, so every variant is its own process.
    python profiles/perf_tma_ab.py            -> gpurun_out/r02_tma_gather_ab.json"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
ROOT = %r
sys.path[:0] = [ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")]
import torch
import kgrec_b200 as K
from kgrec_b200.models.base import device_init
dev = torch.device("cuda:0")
D, NB, B, KN = 100, 256, 1024, 10
res = {}
for E in (100_000, 500_000, 5_000_000):
    with device_init(dev):
        m = K.TransEModel(False, D, E, 500)
    m.grad_mode = "sparse"
    g = torch.Generator().manual_seed(3)
    sets = []
    for _ in range(2):
        n = NB * B
        ph, pt = (torch.randint(0, E, (n,), generator=g, dtype=torch.int32).to(dev) for _ in range(2))
        pr = torch.randint(0, 500, (n,), generator=g, dtype=torch.int32).to(dev)
        c = torch.randint(0, E, (n * KN,), generator=g, dtype=torch.int32)
        c = torch.where(torch.rand(n * KN, generator=g) < 0.5, ~c, c).to(dev)
        sets.append((ph, pt, pr, c))
    def step(i):
        ph, pt, pr, c = sets[i %% 2]
        m.zero_grad(set_to_none=True)
        return m.loss_step_corrupt((ph, pt, pr), c, margin=1.0, batch_pos=B)
    for i in range(3): step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20): l = step(i)
    b.record(); torch.cuda.synchronize()
    res[str(E)] = {"ms": a.elapsed_time(b) / 20, "loss0": float(l[0][0])}
    del m
print(json.dumps(res))
''' % ROOT
out = {}
for mode in ("default", "n", "t82", "t83", "t84", "tc2", "tc3", "tg2"):
    env = dict(os.environ)
    env.pop("KGREC_GROUP_STEP", None)
    if mode != "default":
        env["KGREC_GROUP_STEP"] = mode
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    try:
        out[mode] = json.loads(line)
    except Exception:
        out[mode] = {"error": (r.stderr or r.stdout)[-600:]}
    print(mode, out[mode], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_tma_gather_ab.json"), "w"), indent=1)
