#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: scored triples/pairs per second at d=100).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): TransE, d=100, |E|=100k, |R|=500, batches of 1024 positives
with 10 corrupted negatives each.  One STEP = one pass of the hot path over a stream of
`--batches-per-step` such batches: the fused gather -> residual -> L2 reduce -> margin-loss
forward kernel and the sparse-row-gradient backward kernel (a single 11 264-triple batch is
~14 MB = 2 us of HBM time, i.e. launch-latency bound; SURVEY.md 8d asks for the steady-state
stream).  `value` = scored triples / s with indices resident in HBM; `e2e` = the same through
the module API with pinned host index buffers copied in and the per-batch losses copied out
inside the timed region.  N > 1: independent replicas, one per GPU (the training path does not
shard), value = sum over ranks / max-over-ranks time.

The line also carries `eval`: full-catalog TransE evaluation (every query against every entity,
on-chip top-10) with the entity table row-sharded over the N GPUs and one NCCL all-gather of the
per-shard top-K -- the path's only collective.

--impl reference: the reference's own CPU path for the same step -- its op sequence restated in
torch-CPU (oracle/torch_port.py; the Python reference cannot travel to the GPU box) -- timed on
the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "joint-kg-recommender_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

D, N_ENT, N_REL, BATCH, K_NEG = 100, 100_000, 500, 1024, 10
METRIC = "scored (h,r,t) triples/s, TransE d=100 fused forward+loss+backward"
ROW = 4 * D
# algorithmic bytes per scored triple, independent-triple accounting
FWD_BYTES = 3 * ROW + 3 * 4 + 4                 # 3 rows + 3 int32 ids + 1 score
BWD_BYTES = 2 * 3 * ROW + 16                    # re-gather 3 rows + write 3 gradient rows + ids/score
# fused group accounting (a negative shares 2 of its 3 rows with its positive)
FWD_GROUP_BYTES = ((3 + K_NEG) * ROW + (3 + K_NEG) * 4 + (1 + K_NEG) * 4) / (1 + K_NEG)     # 481 B
BWD_GROUP_BYTES = (2 * (3 + K_NEG) * ROW + (3 + K_NEG) * 4 + (1 + K_NEG) * 4) / (1 + K_NEG)  # 954 B


def ncu_traffic(kernel, batches_per_step):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (same workload only)."""
    if batches_per_step != 256:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[kernel]["bytes"]
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 7 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 7 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_indices(torch, gen, n_batches):
    """Synthetic positives + corrupt-head/tail negatives (utils/data.py:12-18), int32, host pinned."""
    n_pos = n_batches * BATCH
    ph = torch.randint(0, N_ENT, (n_pos,), generator=gen, dtype=torch.int32)
    pt = torch.randint(0, N_ENT, (n_pos,), generator=gen, dtype=torch.int32)
    pr = torch.randint(0, N_REL, (n_pos,), generator=gen, dtype=torch.int32)
    nh = ph.repeat_interleave(K_NEG)
    nt = pt.repeat_interleave(K_NEG)
    nr = pr.repeat_interleave(K_NEG)
    corrupt = torch.randint(0, N_ENT, (n_pos * K_NEG,), generator=gen, dtype=torch.int32)
    head = torch.rand(n_pos * K_NEG, generator=gen) < 0.5
    nh = torch.where(head, corrupt, nh)
    nt = torch.where(head, nt, corrupt)
    cfmt = torch.where(head, ~corrupt, corrupt)              # group-compact format: sign bit = head replaced
    return [x.contiguous() for x in (ph, pt, pr, nh, nt, nr, cfmt)]


def pick_threads(torch, step):
    """The reference path is O(table) per step and scales badly past a few dozen threads:
    time two steps at a handful of thread counts and keep the fastest (reported as `cores`)."""
    cores = os.cpu_count() or 1
    best, best_t = cores, None
    for nt in sorted({min(cores, x) for x in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(nt)
        step()
        t0 = time.perf_counter()
        step(); step()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    torch.set_num_threads(best)
    return best, cores


def run_reference(args):
    """CPU arm: the reference's op sequence (torch-CPU port) on a bounded sample of the workload."""
    import torch
    from oracle import torch_port as TP
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1)
    model = TP.TransPort(False, D, N_ENT, N_REL, with_norm=False)
    sample_batches = 1                                   # one 1024+10240 batch per step (dense-grad cost ~0.1 s)
    idx = [x.long() for x in make_indices(torch, gen, sample_batches)]
    pos, neg = tuple(idx[:3]), tuple(idx[3:6])
    cores, host_cores = pick_threads(torch, lambda: TP.train_step(model, pos, neg))
    for _ in range(max(1, args.warmup)):
        TP.train_step(model, pos, neg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        TP.train_step(model, pos, neg)
    dt = (time.perf_counter() - t0) / args.steps
    triples = sample_batches * BATCH * (1 + K_NEG)
    val = triples / dt
    sample = ("%d batch(es) of %d pos + %d neg per step, forward+marginLoss+dense backward; best of 8/16/32/64/%d "
              "threads on a %d-core host" % (sample_batches, BATCH, BATCH * K_NEG, host_cores, host_cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "triples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "transe d=100 |E|=100k |R|=500, batch 1024 pos + 10 neg/pos (configs[1])",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": "triples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_leg(torch, seconds=12.0):
    from oracle import torch_port as TP
    gen = torch.Generator().manual_seed(1)
    model = TP.TransPort(False, D, N_ENT, N_REL, with_norm=False)
    idx = [x.long() for x in make_indices(torch, gen, 1)]
    pos, neg = tuple(idx[:3]), tuple(idx[3:6])
    cores, host_cores = pick_threads(torch, lambda: TP.train_step(model, pos, neg))
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds and n < 200:
        TP.train_step(model, pos, neg)
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": BATCH * (1 + K_NEG) / dt, "unit": "triples/s", "cores": cores, "kind": "port",
            "sample": "%d steps of one 1024 pos + 10240 neg batch: forward + marginLoss + dense-gradient backward "
                      "(oracle/torch_port.py, the reference's op sequence on torch-CPU); fastest of 8/16/32/64/%d "
                      "threads on the %d-core host" % (n, host_cores, host_cores)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import kgrec_b200 as K
    from kgrec_b200 import evaluation as KE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL's version / debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1234 + rank)
    nb = args.batches_per_step
    n_pos = nb * BATCH
    n_tri = n_pos * (1 + K_NEG)

    model = K.TransEModel(False, D, N_ENT, N_REL)
    model.grad_mode = "sparse"
    n_sets = 3                                           # rotate index sets so no step re-reads its ids from L2
    host_sets = [[x.pin_memory() for x in make_indices(torch, gen, nb)] for _ in range(n_sets)]
    dev_sets = [[x.to(dev) for x in hs] for hs in host_sets]
    loss_host = torch.empty(nb, dtype=torch.float32).pin_memory()

    def step_device(s, ev=None, mode="step"):
        ix = dev_sets[s % n_sets]
        model.zero_grad(set_to_none=True)
        if ev: ev[0].record()
        if mode == "step":      # forward + margin loss + backward in one kernel, group-compact negatives
            loss, _, _ = model.loss_step_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=BATCH)
            if ev: ev[1].record(); ev[2].record()
            return loss
        if mode == "generic":   # negatives as full (nh, nt, nr) triples, the reference drivers' format
            loss, _, _ = model.rank_loss(tuple(ix[:3]), tuple(ix[3:6]), margin=1.0, batch_pos=BATCH)
        else:                   # group-compact negatives, separate forward and autograd backward
            loss, _, _ = model.rank_loss_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=BATCH)
        if ev: ev[1].record()
        loss.sum().backward()
        if ev: ev[2].record()
        return loss

    def step_e2e(s):
        hs = host_sets[s % n_sets]
        ix = [hs[i].to(dev, non_blocking=True) for i in (0, 1, 2, 6)]
        model.zero_grad(set_to_none=True)
        loss, _, _ = model.loss_step_corrupt(tuple(ix[:3]), ix[3], margin=1.0, batch_pos=BATCH)
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller reads the losses
        return loss_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ---------------------------------------------------------
    sampler = ClockSampler(local)          # samples from the warm-up to the end of the e2e leg (all under load)
    if rank == 0:
        sampler.start()
    for s in range(max(args.warmup, 50)):  # >= 50 ms of load so the clock record has samples
        step_device(s)
    launches0 = model.kernel_launches
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for s in range(args.steps):
        step_device(s, evs[s])
    t_end.record()
    barrier()
    total_ms = max_over_ranks(t_begin.elapsed_time(t_end))
    step_kernel_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps   # k_group_step + k_batch_loss + glue
    launches = model.kernel_launches - launches0
    ms_per_step = total_ms / args.steps
    value = world * n_tri / (ms_per_step * 1e-3)

    # ---- end-to-end through the module API with host buffers ----------------------------
    # ids start in pinned host memory; the package's DevicePrefetcher copies step i+1 on a side
    # stream while step i runs; every step's per-batch losses are copied back to the host and the
    # loop ends with a stream sync, so all H2D / D2H traffic is inside the timed region.
    from kgrec_b200.data import DevicePrefetcher

    def host_batches(n):
        for s in range(n):
            hs = host_sets[s % n_sets]
            yield [hs[0], hs[1], hs[2], hs[6]]

    def run_e2e(n):
        for ix in DevicePrefetcher(host_batches(n), dev):
            model.zero_grad(set_to_none=True)
            loss, _, _ = model.loss_step_corrupt(tuple(ix[:3]), ix[3], margin=1.0, batch_pos=BATCH)
            loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller reads the losses
    run_e2e(args.warmup)
    barrier()
    e2e_runs = []
    for _ in range(3):            # K steps, three times; the median run is reported (a 10 ms window is
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # sensitive to one host hiccup)
        e0.record()
        run_e2e(args.steps)
        e1.record()
        barrier()
        e2e_runs.append(max_over_ranks(e0.elapsed_time(e1)) / args.steps)
    e2e_ms = sorted(e2e_runs)[1]
    # the same without overlap (copy, compute, read back, one step at a time)
    for s in range(2):
        step_e2e(s)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for s in range(args.steps):
        step_e2e(s)
    f1.record()
    barrier()
    e2e_sync_ms = max_over_ranks(f0.elapsed_time(f1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(host_sets[0][i].numel() * host_sets[0][i].element_size() for i in (0, 1, 2, 6))
    e2e = {"value": world * n_tri / (e2e_ms * 1e-3), "unit": "triples/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": nb * 4, "ms_per_step": e2e_ms,
           "pipeline": "DevicePrefetcher: H2D of step i+1 overlaps step i; losses read back every step; "
                       "median of 3 runs of K steps (ms each: %s)" % ", ".join("%.3f" % x for x in e2e_runs),
           "unpipelined_value": world * n_tri / (e2e_sync_ms * 1e-3)}

    # ---- the same work as separate forward / autograd-backward kernels, in both negative formats
    split = {}
    for mode in ("split", "generic"):
        for s in range(2):
            step_device(s, mode=mode)
        torch.cuda.synchronize()
        gevs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        for s in range(args.steps):
            step_device(s, gevs[s], mode=mode)
        torch.cuda.synchronize()
        split[mode] = (sum(e[0].elapsed_time(e[1]) for e in gevs) / args.steps,
                       sum(e[1].elapsed_time(e[2]) for e in gevs) / args.steps)
    fwd_ms, bwd_ms = split["split"]
    gen_fwd_ms, gen_bwd_ms = split["generic"]

    # ---- complete training step: fused step into accumulators + global-norm clip + sparse-row Adagrad
    from kgrec_b200.optim import SparseRowOptimizer
    omodel = K.TransEModel(False, D, N_ENT, N_REL)
    opt = SparseRowOptimizer(omodel, optimizer_type="Adagrad", lr=0.01, clip=5.0)
    for s in range(3):
        ix = dev_sets[s % n_sets]
        opt.step_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=BATCH)
    torch.cuda.synchronize()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for s in range(args.steps):
        ix = dev_sets[s % n_sets]
        opt.step_corrupt(tuple(ix[:3]), ix[6], margin=1.0, batch_pos=BATCH)
    o1.record()
    torch.cuda.synchronize()
    opt_ms = o0.elapsed_time(o1) / args.steps
    # ... and with the negatives drawn on the device as well (filtered against the positives)
    from kgrec_b200.sampling import TripleNegativeSampler
    kn = torch.stack([dev_sets[0][0], dev_sets[0][1], dev_sets[0][2]], dim=1).long().cpu()
    sampler = TripleNegativeSampler(N_ENT, N_REL, kn, device=dev)
    for s in range(3):
        ix = dev_sets[s % n_sets]
        opt.step_corrupt(tuple(ix[:3]), sampler.sample(tuple(ix[:3]), K_NEG, seed=s), margin=1.0, batch_pos=BATCH)
    torch.cuda.synchronize()
    o0.record()
    for s in range(args.steps):
        ix = dev_sets[s % n_sets]
        opt.step_corrupt(tuple(ix[:3]), sampler.sample(tuple(ix[:3]), K_NEG, seed=100 + s), margin=1.0, batch_pos=BATCH)
    o1.record()
    torch.cuda.synchronize()
    loop_ms = o0.elapsed_time(o1) / args.steps
    del omodel, opt, sampler

    # ---- single-batch latency (the reference's actual training shape) --------------------
    small = [x[:BATCH * (1 if i < 3 else K_NEG)].contiguous() for i, x in enumerate(dev_sets[0])]
    for _ in range(5):
        model.loss_step_corrupt(tuple(small[:3]), small[6], margin=1.0)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(50):
        model.zero_grad(set_to_none=True)
        model.loss_step_corrupt(tuple(small[:3]), small[6], margin=1.0)
    s1.record()
    torch.cuda.synchronize()
    single_us = s0.elapsed_time(s1) * 1e3 / 50

    # ---- full-catalog evaluation, catalog sharded over the N GPUs -------------------------
    ev = None
    if not args.no_eval:
        nq, n_cat = args.eval_queries, args.eval_entities
        torch.manual_seed(7)                                  # the same table on every rank
        emodel = K.TransEModel(False, D, n_cat, N_REL)
        lo, hi = KE.shard_bounds(n_cat, world, rank)
        shard = emodel.ent_embeddings.weight.detach()[lo:hi].contiguous()
        qg = torch.Generator().manual_seed(99)
        qh = torch.randint(0, n_cat, (nq,), generator=qg).to(dev)
        qr = torch.randint(0, N_REL, (nq,), generator=qg).to(dev)

        def eval_pass():
            keys = emodel.topk("tail", qh, qr, k=10, catalog=shard, id_base=lo)
            return KE.sharded_topk(keys) if world > 1 else keys
        for _ in range(2):
            eval_pass()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        reps = 3
        for _ in range(reps):
            eval_pass()
        a1.record()
        barrier()
        ems = max_over_ranks(a0.elapsed_time(a1)) / reps
        ev = {"metric": "scored (query,entity) pairs/s, TransE L2 d=100 full-catalog top-10", "value": nq * n_cat / (ems * 1e-3),
              "unit": "pairs/s", "ms": ems, "queries": nq, "catalog_rows": n_cat, "catalog_sharding": "rows / %d GPUs" % world,
              "collective": "1 NCCL all-gather of [nq,10] uint64 keys per pass" if world > 1 else "none (1 GPU)",
              "bound": "fp32 pipe: 2 lane-ops per (pair, dim) on 148 SM x 128 lanes x 1.965 GHz",
              "fp32_bound_pairs_per_s": 148 * 128 * 1.965e9 / (2 * D) * world,
              "frac_of_fp32_bound": nq * n_cat / (ems * 1e-3) / (148 * 128 * 1.965e9 / (2 * D) * world)}

    # ---- rec-side evaluation (BASELINE configs[2] shape): TUP d=100, P=20, 50k users x 50k items,
    # soft preferences (the shipped transup.sh setting), top-10 per user, on this rank only
    ev_rec = None
    if not args.no_eval and rank == 0:
        torch.manual_seed(11)
        rmodel = K.TransUPModel(False, D, 50_000, 50_000, 20, False)
        qu = torch.arange(0, args.eval_queries, device=dev) % 50_000
        soft_cat = rmodel.soft_catalog()
        for _ in range(2):
            rmodel.topk_items(qu, k=10, soft_catalog=soft_cat)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(3):
            rmodel.topk_items(qu, k=10, soft_catalog=soft_cat)
        r1.record()
        torch.cuda.synchronize()
        rms = r0.elapsed_time(r1) / 3
        ev_rec = {"metric": "scored (user,item) pairs/s, TUP soft d=100 P=20 full-catalog top-10", "value": args.eval_queries * 50_000 / (rms * 1e-3),
                  "unit": "pairs/s", "ms": rms, "users": args.eval_queries, "items": 50_000, "n_gpus": 1,
                  "note": "augmented item rows built once (soft_catalog), user rows per call"}
        del rmodel, soft_cat
    # ---- rec-side training step (BASELINE configs[2] / [3] shapes): 256 batches x (1024 positives + 1 negative
    # each), forward + BPR loss + backward in one pass of the tile kernel (kgrec_rank_loss_step), this rank only
    tr_rec = None
    if not args.no_eval and rank == 0:
        import numpy as np
        tr_rec = {}
        n_pos = 256 * BATCH
        fp32_fma_per_s = 148 * 128 * 1.965e9
        for name, gum, fma in (("tup_st_gumbel", True, 14000), ("ktup_soft", False, 18000)):
            torch.manual_seed(13)
            if name.startswith("tup"):
                tm = K.TransUPModel(False, D, 50_000, 50_000, 20, gum)
            else:
                n_item, n_ent = 50_000, 500_000
                ents = np.random.RandomState(0).permutation(n_ent)[:n_item]
                new_map = {i: (int(ents[i]) if i % 10 < 7 else -1, i) for i in range(n_item)}
                tm = K.jTransUPModel(False, D, 50_000, n_item, n_ent, 20, {i: i for i in range(n_item)}, new_map, False, gum)
            tm.grad_mode = "sparse"
            tg = torch.Generator().manual_seed(5)
            tu, ti, tn = (torch.randint(0, 50_000, (n_pos,), generator=tg, dtype=torch.int32).to(dev) for _ in range(3))
            for _ in range(3):
                tm.zero_grad(set_to_none=True)
                tm.loss_step((tu, ti), (tu, tn), target=-1.0, batch_pos=BATCH)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(5):
                tm.zero_grad(set_to_none=True)
                tm.loss_step((tu, ti), (tu, tn), target=-1.0, batch_pos=BATCH)
            t1.record()
            torch.cuda.synchronize()
            tms = t0.elapsed_time(t1) / 5
            pps = 2 * n_pos / (tms * 1e-3)
            tr_rec[name] = {"metric": "scored (user,item) pairs/s, forward + BPR loss + backward, d=100 P=20", "value": pps,
                            "unit": "pairs/s", "ms": tms, "pairs_per_step": 2 * n_pos, "bound": "fp32 pipe",
                            "fma_per_pair": fma, "frac_of_fp32_bound": pps * fma / fp32_fma_per_s}
            del tm
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    n_fwd, n_bwd = n_tri, n_tri
    fwd_gbs = n_fwd * FWD_GROUP_BYTES / (fwd_ms * 1e-3) / 1e9
    bwd_gbs = n_bwd * BWD_GROUP_BYTES / (bwd_ms * 1e-3) / 1e9
    step_gbs = n_tri * BWD_GROUP_BYTES / (step_kernel_ms * 1e-3) / 1e9
    dom = "k_group_step_e"
    roof = {"bound": "hbm", "kernel": dom, "achieved": step_gbs, "peak": peak,
            "unit": "GB/s", "frac": step_gbs / peak,
            "traffic": ncu_traffic(dom, nb),
            "peak_source": peak_src,
            "bytes_per_triple": BWD_GROUP_BYTES,
            "note": "algorithmic bytes per scored triple in the fused pos + 10 neg group accounting of SURVEY 8d "
                    "((3+K) rows read [+ (3+K) gradient rows written] per 1+K triples); the event interval also "
                    "covers the torch glue around the launch; traffic: ncu capture under profiles/"}
    out = {
        "metric": METRIC, "value": value, "unit": "triples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "transe d=100 |E|=100k |R|=500, batch 1024 pos + 10 neg/pos (configs[1]); "
                               "%d batches per step in one launch: forward + margin loss + sparse-row-gradient "
                               "backward fused in k_group_step_e (COO slot ids written by the same pass); negatives in the group-compact corrupted-id format" % nb,
                   "triples_per_step_per_gpu": n_tri, "index_dtype": "int32", "grad_mode": "sparse slots",
                   "parallelism": "replicas x%d (training path does not shard)" % world,
                   "l2": "inputs larger than L2: per step 14 MB of ids + 1.4 GB of gradient rows stream through the "
                         "126 MB L2; index sets rotate between steps; the 40 MB entity table of configs[1] is "
                         "L2-resident by construction"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roof,
        "kernels": {
            "k_group_step_e": {"ms": step_kernel_ms, "triples_per_s": n_tri / (step_kernel_ms * 1e-3),
                             "algorithmic_GBps": step_gbs, "frac_of_peak": step_gbs / peak,
                             "bytes_per_triple": BWD_GROUP_BYTES},
            "k_group_fwd": {"ms": fwd_ms, "triples_per_s": n_fwd / (fwd_ms * 1e-3), "algorithmic_GBps": fwd_gbs,
                            "frac_of_peak": fwd_gbs / peak, "bytes_per_triple": FWD_GROUP_BYTES},
            "k_group_step_e<backward mode>": {"ms": bwd_ms, "triples_per_s": n_bwd / (bwd_ms * 1e-3), "algorithmic_GBps": bwd_gbs,
                            "frac_of_peak": bwd_gbs / peak, "bytes_per_triple": BWD_GROUP_BYTES},
            "generic_triple_format": {
                "k_rank_loss_fwd": {"ms": gen_fwd_ms, "algorithmic_GBps": n_fwd * FWD_BYTES / (gen_fwd_ms * 1e-3) / 1e9,
                                    "bytes_per_triple": FWD_BYTES},
                "k_score_bwd": {"ms": gen_bwd_ms, "algorithmic_GBps": n_bwd * BWD_BYTES / (gen_bwd_ms * 1e-3) / 1e9,
                                "bytes_per_triple": BWD_BYTES},
                "triples_per_s": n_tri / ((gen_fwd_ms + gen_bwd_ms) * 1e-3)},
        },
        "single_batch_latency_us": single_us,
        "full_train_step": {"what": "k_group_step_e (dense accumulate) + clip_grad_norm(5) + sparse-row Adagrad update of the "
                                    "touched rows (kgrec_b200.optim.SparseRowOptimizer), %d batches per step" % nb,
                            "ms": opt_ms, "triples_per_s": n_tri / (opt_ms * 1e-3),
                            "with_device_negative_sampling_ms": loop_ms,
                            "with_device_negative_sampling_triples_per_s": n_tri / (loop_ms * 1e-3)},
        "eval": ev,
        "eval_rec": ev_rec,
        "train_rec": tr_rec,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_leg(torch)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batches-per-step", type=int, default=256)
    ap.add_argument("--eval-queries", type=int, default=4096)
    ap.add_argument("--eval-entities", type=int, default=1_000_000)
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
