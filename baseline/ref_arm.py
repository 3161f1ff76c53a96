"""The reference's own CPU path, timed: what ``bench.py --impl reference`` and the
``cpu_baseline`` key of the GPU arm report.

The model classes are the UNMODIFIED reference classes imported from ``baseline/_ref``
(``baseline/make_ref.py``: a copy of the reference whose only edits are the four torch>=0.4
driver patches of SURVEY.md 8c -- none of them touches transE/transH/transUP forward code; the
jTransUP ``paddingItems(...tolist())`` patch only restores its dict lookup).  They run on the host
cores (``CUDA_VISIBLE_DEVICES`` is irrelevant here: ``to_gpu`` is bypassed by constructing while
``jTransUP.utils.misc.USE_CUDA`` is forced False), with ``torch.set_num_threads`` chosen as the
fastest of a few counts and reported.

Regions follow BASELINE.md section 4: (i) ``model(pos) + model(neg) + loss`` forward,
(ii) forward + backward (dense ``[rows, d]`` gradients, as the reference's autograd produces),
(iii) ``evaluate*`` score-matrix production on a memory-feasible query slice.
Throughput unit: scored triples (or (user, item) pairs) per second; a batch of P positives with
K negatives each counts P * (1 + K) scored triples for BOTH arms, although the reference has to
score every positive K times to pair it with its negatives (``pos.repeat_interleave(K)``,
SURVEY 8d cfg#2 note).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref  # noqa: E402

_ref = {}


def available():
    return make_ref.available()


def load():
    """Import the reference classes from baseline/_ref, pinned to the host."""
    if _ref:
        return _ref
    import torch
    for p in reversed(make_ref.env_paths()):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")
    import gflags  # noqa: F401   (shim: absl.flags + numpy.asfarray)
    import jTransUP.utils.misc as misc
    misc.USE_CUDA = False                    # latched at import (misc.py:11); the baseline is the CPU path
    from jTransUP.models.transE import TransEModel
    from jTransUP.models.transH import TransHModel
    from jTransUP.models.transR import TransRModel
    from jTransUP.models.transUP import TransUPModel
    from jTransUP.models.jTransUP import jTransUPModel
    from jTransUP.utils import loss as L
    for mod in ("transE", "transH", "transR", "transUP", "jTransUP"):
        m = sys.modules["jTransUP.models." + mod]
        if hasattr(m, "to_gpu"):
            m.to_gpu = lambda x: x           # the modules bound misc.to_gpu at import
    L.to_gpu = lambda x: x
    _ref.update(torch=torch, TransE=TransEModel, TransH=TransHModel, TransR=TransRModel, TUP=TransUPModel,
                KTUP=jTransUPModel, loss=L)
    return _ref


def pick_threads(torch, fn, counts=None):
    """The reference path is O(table) per step and scales badly past a few dozen threads: time fn at
    a handful of thread counts and keep the fastest (reported as `cores`)."""
    cores = os.cpu_count() or 1
    best, best_t = cores, None
    for _ in range(2):          # the first two passes over fresh tables are far off the steady state: not charged to a count
        fn()
    for nt in sorted({min(cores, x) for x in (counts or (8, 16, 32, 64, cores))}):
        torch.set_num_threads(nt)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    torch.set_num_threads(best)
    return best, cores


def tried_counts(cores, counts=None):
    """The thread counts pick_threads times on a host with `cores` cores, as text."""
    return "/".join(str(x) for x in sorted({min(cores, x) for x in (counts or (8, 16, 32, 64, cores))}))


def _best_of(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


# ---------------------------------------------------------------------------------------------------------
# KG: TransE / TransH training step of the reference driver (knowledge_representation.py:187-207)
# ---------------------------------------------------------------------------------------------------------
class KGStep:
    """model(pos) + model(neg) + marginLoss [+ backward], 1024 positives x K negatives per batch."""

    def __init__(self, model_type="transe", d=100, n_ent=100_000, n_rel=500, batch=1024, k_neg=10, l1=False, seed=0):
        R = load()
        torch = R["torch"]
        torch.manual_seed(seed)
        cls = {"transe": R["TransE"], "transh": R["TransH"]}[model_type]
        self.model = cls(l1, d, n_ent, n_rel)
        self.torch, self.L, self.batch, self.k_neg, self.n_ent, self.n_rel = torch, R["loss"], batch, k_neg, n_ent, n_rel
        self.kind = "reference"
        self.make_batch(seed)

    def make_batch(self, seed):
        torch = self.torch
        g = torch.Generator().manual_seed(seed + 1)
        B, K = self.batch, self.k_neg
        ph = torch.randint(0, self.n_ent, (B,), generator=g)
        pt = torch.randint(0, self.n_ent, (B,), generator=g)
        pr = torch.randint(0, self.n_rel, (B,), generator=g)
        c = torch.randint(0, self.n_ent, (B * K,), generator=g)
        head = torch.rand(B * K, generator=g) < 0.5
        rep = lambda x: x.repeat_interleave(K)           # noqa: E731
        self.pos = (rep(ph), rep(pt), rep(pr))           # the reference pairs score vectors element by element
        self.neg = (torch.where(head, c, rep(ph)), torch.where(head, rep(pt), c), rep(pr))
        self.units = B * (1 + K)

    def forward(self):
        with self.torch.no_grad():
            return self.L.marginLoss()(self.model(*self.pos), self.model(*self.neg), 1.0)

    def forward_backward(self):
        self.model.zero_grad()
        loss = self.L.marginLoss()(self.model(*self.pos), self.model(*self.neg), 1.0)
        loss.backward()
        return loss

    def evaluate(self, n_queries=16, side="tail"):
        torch = self.torch
        q = torch.arange(n_queries) % self.n_ent
        r = torch.arange(n_queries) % self.n_rel
        with torch.no_grad():
            return (self.model.evaluateTail if side == "tail" else self.model.evaluateHead)(q, r)


class RecStep:
    """TUP / KTUP: model(u, pi) + model(u, ni) + bprLoss [+ backward] (item_recommendation.py:171-175,
    knowledgable_recommendation.py:337-341), one negative per positive as the reference samples them."""

    def __init__(self, ktup=False, d=100, n_user=50_000, n_item=50_000, n_pref=20, n_ent=0, gumbel=True, batch=1024, seed=0):
        R = load()
        torch = R["torch"]
        torch.manual_seed(seed)
        self.torch, self.L, self.batch, self.ktup = torch, R["loss"], batch, ktup
        if ktup:
            ents = np.random.RandomState(seed).permutation(n_ent)[:n_item]
            new_map = {i: ((int(ents[i]) if i % 10 < 7 else -1), i) for i in range(n_item)}
            self.model = R["KTUP"](False, d, n_user, n_item, n_ent, n_pref, {i: i for i in range(n_item)}, new_map, False, gumbel)
        else:
            self.model = R["TUP"](False, d, n_user, n_item, n_pref, gumbel)
        g = torch.Generator().manual_seed(seed + 1)
        self.u = torch.randint(0, n_user, (batch,), generator=g)
        self.pi = torch.randint(0, n_item, (batch,), generator=g)
        self.ni = torch.randint(0, n_item, (batch,), generator=g)
        self.n_user, self.n_ent, self.n_pref = n_user, n_ent, n_pref
        self.units = 2 * batch
        if ktup:
            self.kg = (torch.randint(0, n_ent, (batch,), generator=g), torch.randint(0, n_ent, (batch,), generator=g),
                       torch.randint(0, n_pref, (batch,), generator=g), torch.randint(0, n_ent, (batch,), generator=g))

    def _scores(self):
        if self.ktup:
            return self.model((self.u, self.pi), None, is_rec=True), self.model((self.u, self.ni), None, is_rec=True)
        return self.model(self.u, self.pi), self.model(self.u, self.ni)

    def forward(self):
        with self.torch.no_grad():
            return self.L.bprLoss(*self._scores(), target=-1)

    def forward_backward(self):
        self.model.zero_grad()
        loss = self.L.bprLoss(*self._scores(), target=-1)
        loss.backward()
        return loss

    def kg_forward_backward(self):
        """KTUP's KG branch (knowledgable_recommendation.py:368-374): TransH on ent / rel / norm, tail-corrupted."""
        h, t, r, c = self.kg
        self.model.zero_grad()
        loss = self.L.marginLoss()(self.model(None, (h, t, r), is_rec=False), self.model(None, (h, c, r), is_rec=False), 1.0)
        loss.backward()
        return loss

    def evaluate(self, n_queries=8):
        q = self.torch.arange(n_queries) % self.n_user
        with self.torch.no_grad():
            return self.model.evaluateRec(q) if self.ktup else self.model.evaluate(q)


class PortKGStep:
    """Fallback when baseline/_ref is absent (a checkout without the reference): the oracle's torch-CPU
    restatement of the same op sequence (oracle/torch_port.py)."""

    def __init__(self, d=100, n_ent=100_000, n_rel=500, batch=1024, k_neg=10, seed=0):
        import torch
        sys.path.insert(0, os.path.dirname(HERE))
        from oracle import torch_port as TP
        torch.manual_seed(seed)
        self.torch, self.TP = torch, TP
        self.model = TP.TransPort(False, d, n_ent, n_rel, with_norm=False)
        self.batch, self.k_neg, self.n_ent, self.n_rel = batch, k_neg, n_ent, n_rel
        self.kind = "port"
        KGStep.make_batch(self, seed)

    def forward_backward(self):
        return self.TP.train_step(self.model, self.pos, self.neg)


def headline_step(seed=0):
    """BASELINE.json configs[1] (TransE d=100, |E|=100k, |R|=500, 1024 pos + 10 neg/pos), region (ii)."""
    return KGStep("transe", seed=seed) if available() else PortKGStep(seed=seed)


def time_headline(steps, warmup, batches_per_step, budget_s=None):
    """K steps of `batches_per_step` reference batches each (forward + marginLoss + dense backward)."""
    st = headline_step()
    torch = st.torch
    cores, host = pick_threads(torch, st.forward_backward)

    def step():
        for _ in range(batches_per_step):
            st.forward_backward()
    for _ in range(max(1, warmup)):
        step()
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        step()
        n += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / n
    sample = ("%d step(s) of %d batch(es) of 1024 pos + 10240 neg: model(pos.repeat_interleave(10)) + model(neg) + marginLoss "
              "+ backward (dense [rows, d] gradients) through %s; fastest of %s threads on the %d-core host"
              % (n, batches_per_step, "the unmodified reference classes (baseline/_ref)" if st.kind == "reference"
                 else "oracle/torch_port.py (baseline/_ref absent)", tried_counts(host), host))
    return {"value": batches_per_step * st.units / dt, "unit": "triples/s", "cores": cores, "kind": st.kind,
            "sample": sample, "ms_per_step": dt * 1e3, "steps_run": n}


def regions(reps=2, cfg5=True):
    """BASELINE.md section 4 regions for configs[1..4]; every value in scored units per second."""
    if not available():
        return None
    R = load()
    torch = R["torch"]
    out = {}

    def rate(units, fn):
        # two warm-up calls: the reference's first TWO passes over fresh tables run 10-80x slower than its steady state
        # (first-touch of the [rows, d] temporaries, thread-pool start); best of >= 2 timed calls after that
        return units / _best_of(fn, reps=max(2, reps), warm=2)
    # configs[1]: TransE d=100, 100k entities
    k = KGStep("transe")
    out["cfg2_transe_forward"] = rate(k.units, k.forward)
    out["cfg2_transe_forward_backward"] = rate(k.units, k.forward_backward)
    out["cfg2_transe_evaluateTail_16q_x_100k"] = rate(16 * k.n_ent, lambda: k.evaluate(16))
    del k
    # configs[2]: TUP d=100, P=20, ST-Gumbel, 50k users x 50k items
    t = RecStep(ktup=False, gumbel=True)
    out["cfg3_tup_gumbel_forward"] = rate(t.units, t.forward)
    out["cfg3_tup_gumbel_forward_backward"] = rate(t.units, t.forward_backward)
    out["cfg3_tup_gumbel_evaluate_8u_x_50k"] = rate(8 * 50_000, lambda: t.evaluate(8))
    del t
    # configs[3]: KTUP, ml1m-scale rec (6040 x 3706) + 500k entities, R = P = 20
    j = RecStep(ktup=True, n_user=6040, n_item=3706, n_pref=20, n_ent=500_000, gumbel=False)
    out["cfg4_ktup_rec_forward_backward"] = rate(j.units, j.forward_backward)
    out["cfg4_ktup_kg_forward_backward"] = rate(2 * j.batch, j.kg_forward_backward)
    out["cfg4_ktup_evaluateRec_8u_x_3706"] = rate(8 * 3706, lambda: j.evaluate(8))
    del j
    if cfg5:
        # configs[4] slices: d=128; the reference materialises [B, N, d], so 8 queries x a 500k-row slice
        e = KGStep("transe", d=128, n_ent=500_000, n_rel=500)
        out["cfg5_transe_evaluateTail_8q_x_500k_d128"] = rate(8 * 500_000, lambda: e.evaluate(8))
        del e
        u = RecStep(ktup=False, d=128, n_user=100_000, n_item=100_000, gumbel=False)
        out["cfg5_tup_soft_evaluate_4u_x_100k_d128"] = rate(4 * 100_000, lambda: u.evaluate(4))
        del u
    out["threads"] = torch.get_num_threads()
    return out
