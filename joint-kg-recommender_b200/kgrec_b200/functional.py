"""torch.autograd glue over the C ABI (include/kgrec_b200.h).

Everything here runs the CUDA kernels; tensors on the CPU are rejected with a
RuntimeError (there is no CPU path).  PyTorch supplies device memory, the
current stream and autograd bookkeeping -- nothing else.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import Grads, Tables

_TABLE_FIELDS = ("ent", "rel", "norm", "proj", "user", "item", "pref", "pref_norm")
# which tables each model reads, in the order autograd sees them
MODEL_TABLES = {
    _lib.TRANSE: ("ent", "rel"),
    _lib.TRANSH: ("ent", "rel", "norm"),
    _lib.TRANSR: ("ent", "rel", "proj"),
    _lib.TUP: ("user", "item", "pref", "pref_norm"),
    _lib.KTUP: ("user", "item", "ent", "rel", "norm", "pref", "pref_norm"),
}
# tables whose gradient is a gathered-row ("slot") gradient; the others are small dense tables
_SLOT_TABLES = {
    _lib.TRANSE: ("ent", "rel"),
    _lib.TRANSH: ("ent", "rel", "norm"),
    _lib.TRANSR: ("ent", "rel"),
    _lib.TUP: ("user", "item"),
    _lib.KTUP: ("user", "item", "ent"),
}


# kernels of this library enqueued so far (forward, backward and evaluation paths)
LAUNCHES = [0]


def count_launches(n):
    LAUNCHES[0] += n


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check_table(name, w):
    if not w.is_cuda:
        raise RuntimeError("kgrec_b200: table '%s' is on %s; the engine has no CPU path -- move the module "
                           "to a CUDA device" % (name, w.device))
    if w.dtype != torch.float32 or not w.is_contiguous():
        raise RuntimeError("kgrec_b200: table '%s' must be contiguous float32" % name)


def make_tables(weights, dim, l1, use_gumbel=False, item2ent=None):
    """Build the kgrec_tables struct from a {name: tensor} dict of [rows, d] tables."""
    t = Tables()
    t.dim = dim
    t.ld = dim
    t.l1 = int(bool(l1))
    t.use_gumbel = int(bool(use_gumbel))
    for name in _TABLE_FIELDS:
        w = weights.get(name)
        if w is None:
            continue
        _check_table(name, w)
        setattr(t, name, w.data_ptr())
    for name, field in (("ent", "n_ent"), ("rel", "n_rel"), ("user", "n_user"), ("item", "n_item")):
        if weights.get(name) is not None:
            setattr(t, field, weights[name].shape[0])
    if weights.get("pref") is not None:
        t.n_pref = weights["pref"].shape[0]
    if item2ent is not None:
        if item2ent.dtype != torch.int32 or not item2ent.is_cuda:
            raise RuntimeError("kgrec_b200: item2ent must be an int32 CUDA tensor")
        t.item2ent = item2ent.data_ptr()
    return t


def as_index(x, device):
    """Index arrays arrive as the drivers build them (LongTensor, maybe on the host)."""
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    if x.dtype not in (torch.int64, torch.int32):
        x = x.long()
    if x.device != device:
        x = x.to(device, non_blocking=True)
    return x.contiguous().view(-1)


def _idx_bytes(*xs):
    sizes = {x.element_size() for x in xs if x is not None}
    if len(sizes) != 1:
        raise RuntimeError("kgrec_b200: index arrays must share one integer width")
    return sizes.pop()


def _alloc_grads(model, weights, n, mode, dim):
    """Gradient buffers for one backward call.  Returns (Grads struct, {name: tensor})."""
    dev = weights[MODEL_TABLES[model][0]].device
    g = Grads()
    g.mode = 1 if mode == "dense" else 0
    bufs = {}
    slots = {"ent": 2 * n if model != _lib.KTUP else n, "rel": n, "norm": n, "user": n, "item": n}
    for name in MODEL_TABLES[model]:
        w = weights[name]
        if model == _lib.KTUP and name in ("rel", "norm"):
            continue                       # shares the pref / pref_norm gradient
        if name in _SLOT_TABLES[model] and mode != "dense":
            buf = torch.empty((slots[name], dim), dtype=torch.float32, device=dev)
        else:
            buf = torch.zeros_like(w)
        bufs[name] = buf
        setattr(g, name, buf.data_ptr())
    return g, bufs


def _finish_grads(model, weights, bufs, idx, mode, needs):
    """Turn kernel outputs into what autograd returns for each table."""
    out = {}
    for name in MODEL_TABLES[model]:
        if not needs.get(name, False):
            out[name] = None
            continue
        if model == _lib.KTUP and name in ("rel", "norm"):
            out[name] = bufs["pref" if name == "rel" else "pref_norm"]
            continue
        buf = bufs[name]
        if name in _SLOT_TABLES[model] and mode != "dense":
            w = weights[name]
            out[name] = torch.sparse_coo_tensor(idx[name].view(1, -1), buf, size=tuple(w.shape), check_invariants=False)
        else:
            out[name] = buf
    return out


class _Ctx:
    """Per-call constants shared by forward and backward."""

    def __init__(self, model, dim, l1, use_gumbel, item2ent, grad_mode, seed):
        self.model, self.dim, self.l1, self.use_gumbel = model, dim, l1, use_gumbel
        self.item2ent, self.grad_mode, self.seed = item2ent, grad_mode, seed


def _slot_indices(model, a, b, c, item2ent):
    if model in (_lib.TRANSE, _lib.TRANSH, _lib.TRANSR):
        idx = {"ent": torch.cat([a, b]).long(), "rel": c.long()}
        if model == _lib.TRANSH:
            idx["norm"] = idx["rel"]
        return idx
    idx = {"user": a.long(), "item": b.long()}
    if model == _lib.KTUP:
        idx["ent"] = item2ent[b.long()].long()
    return idx


class ScoreFunction(torch.autograd.Function):
    """scores = model(a, b, c): kgrec_score_fwd / kgrec_score_bwd."""

    @staticmethod
    def forward(ctx, cfg, a, b, c, gumbel_u, status, *tables):
        names = MODEL_TABLES[cfg.model]
        weights = dict(zip(names, tables))
        T = make_tables(weights, cfg.dim, cfg.l1, cfg.use_gumbel, cfg.item2ent)
        n = a.numel()
        scores = torch.empty(n, dtype=torch.float32, device=a.device)
        lib = _lib.load()
        _lib.check(lib.kgrec_score_fwd(C.byref(T), cfg.model, _ptr(a), _ptr(b), _ptr(c), _idx_bytes(a, b, c), n,
                                       _ptr(gumbel_u), cfg.seed, _ptr(scores), _ptr(status), _stream()))
        count_launches(1)
        ctx.cfg = cfg
        ctx.idx = (a, b, c, gumbel_u)
        ctx.save_for_backward(*tables)
        return scores

    @staticmethod
    def backward(ctx, grad_scores):
        cfg = ctx.cfg
        a, b, c, gumbel_u = ctx.idx
        tables = ctx.saved_tensors
        names = MODEL_TABLES[cfg.model]
        weights = dict(zip(names, tables))
        needs = dict(zip(names, ctx.needs_input_grad[6:]))
        n = a.numel()
        T = make_tables(weights, cfg.dim, cfg.l1, cfg.use_gumbel, cfg.item2ent)
        G, bufs = _alloc_grads(cfg.model, weights, n, cfg.grad_mode, cfg.dim)
        gs = grad_scores.contiguous().float()
        lib = _lib.load()
        _lib.check(lib.kgrec_score_bwd(C.byref(T), cfg.model, _ptr(a), _ptr(b), _ptr(c), _idx_bytes(a, b, c), n,
                                       _ptr(gumbel_u), cfg.seed, _ptr(gs), C.byref(G), _stream()))
        count_launches(1)
        idx = _slot_indices(cfg.model, a, b, c, cfg.item2ent) if cfg.grad_mode != "dense" else {}
        out = _finish_grads(cfg.model, weights, bufs, idx, cfg.grad_mode, needs)
        return (None, None, None, None, None, None) + tuple(out[nm] for nm in names)


class RankLossFunction(torch.autograd.Function):
    """loss[b], pos, neg = fused pos + K-negative scoring and margin / BPR loss."""

    @staticmethod
    def forward(ctx, cfg, pos, neg, n_neg, batch_pos, loss_kind, param, gumbel_u, status, *tables):
        names = MODEL_TABLES[cfg.model]
        weights = dict(zip(names, tables))
        T = make_tables(weights, cfg.dim, cfg.l1, cfg.use_gumbel, cfg.item2ent)
        pa, pb, pc = pos
        na, nb, nc = neg
        n_pos = pa.numel()
        dev = pa.device
        lib = _lib.load()
        pos_s = torch.empty(n_pos, dtype=torch.float32, device=dev)
        neg_s = torch.empty(n_pos * n_neg, dtype=torch.float32, device=dev)
        n_batches = (n_pos + batch_pos - 1) // batch_pos
        loss = torch.empty(n_batches, dtype=torch.float32, device=dev)
        ws = torch.empty(max(1, n_pos), dtype=torch.float32, device=dev)
        _lib.check(lib.kgrec_rank_loss_fwd(
            C.byref(T), cfg.model, _ptr(pa), _ptr(pb), _ptr(pc), _ptr(na), _ptr(nb), _ptr(nc),
            _idx_bytes(pa, pb, pc, na, nb, nc), n_pos, n_neg, batch_pos, loss_kind, float(param),
            _ptr(gumbel_u), cfg.seed, _ptr(pos_s), _ptr(neg_s), _ptr(loss), _ptr(ws), _ptr(status), _stream()))
        count_launches(2)
        ctx.cfg = cfg
        ctx.args = (pos, neg, n_neg, batch_pos, loss_kind, float(param), gumbel_u, pos_s, neg_s)
        ctx.save_for_backward(*tables)
        ctx.mark_non_differentiable(pos_s, neg_s)
        return loss, pos_s, neg_s

    @staticmethod
    def backward(ctx, grad_loss, _gp, _gn):
        cfg = ctx.cfg
        pos, neg, n_neg, batch_pos, loss_kind, param, gumbel_u, pos_s, neg_s = ctx.args
        pa, pb, pc = pos
        na, nb, nc = neg
        tables = ctx.saved_tensors
        names = MODEL_TABLES[cfg.model]
        weights = dict(zip(names, tables))
        needs = dict(zip(names, ctx.needs_input_grad[9:]))
        n_pos = pa.numel()
        n = n_pos * (1 + n_neg)
        gl = grad_loss.reshape(-1).contiguous().float()     # upstream per loss batch, stays on the device
        T = make_tables(weights, cfg.dim, cfg.l1, cfg.use_gumbel, cfg.item2ent)
        G, bufs = _alloc_grads(cfg.model, weights, n, cfg.grad_mode, cfg.dim)
        lib = _lib.load()
        _lib.check(lib.kgrec_rank_loss_bwd(
            C.byref(T), cfg.model, _ptr(pa), _ptr(pb), _ptr(pc), _ptr(na), _ptr(nb), _ptr(nc),
            _idx_bytes(pa, pb, pc, na, nb, nc), n_pos, n_neg, batch_pos, loss_kind, param,
            _ptr(gumbel_u), cfg.seed, _ptr(pos_s), _ptr(neg_s), 1.0, _ptr(gl), C.byref(G), _stream()))
        count_launches(1)
        idx = {}
        if cfg.grad_mode != "dense":
            pi = _slot_indices(cfg.model, pa, pb, pc, cfg.item2ent)
            ni = _slot_indices(cfg.model, na, nb, nc, cfg.item2ent)
            for k in pi:
                if k == "ent" and cfg.model != _lib.KTUP:
                    # slot order: heads of all n triples, then tails
                    idx[k] = torch.cat([pa, na, pb, nb]).long()
                else:
                    idx[k] = torch.cat([pi[k], ni[k]])
        out = _finish_grads(cfg.model, weights, bufs, idx, cfg.grad_mode, needs)
        return (None,) * 9 + tuple(out[nm] for nm in names)


class CorruptLossFunction(torch.autograd.Function):
    """Fused ranking loss in the group-compact negative format (kgrec_corrupt_loss_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, cfg, pos, corrupt, n_neg, batch_pos, loss_kind, param, status, *tables):
        names = MODEL_TABLES[cfg.model]
        weights = dict(zip(names, tables))
        T = make_tables(weights, cfg.dim, cfg.l1, cfg.use_gumbel, cfg.item2ent)
        ph, pt, pr = pos
        n_pos = ph.numel()
        dev = ph.device
        lib = _lib.load()
        pos_s = torch.empty(n_pos, dtype=torch.float32, device=dev)
        neg_s = torch.empty(n_pos * n_neg, dtype=torch.float32, device=dev)
        loss = torch.empty((n_pos + batch_pos - 1) // batch_pos, dtype=torch.float32, device=dev)
        ws = torch.empty(max(1, n_pos), dtype=torch.float32, device=dev)
        _lib.check(lib.kgrec_corrupt_loss_fwd(
            C.byref(T), cfg.model, _ptr(ph), _ptr(pt), _ptr(pr), _idx_bytes(ph, pt, pr), n_pos, _ptr(corrupt),
            n_neg, batch_pos, loss_kind, float(param), _ptr(pos_s), _ptr(neg_s), _ptr(loss), _ptr(ws),
            _ptr(status), _stream()))
        count_launches(2)
        ctx.cfg = cfg
        ctx.args = (pos, corrupt, n_neg, batch_pos, loss_kind, float(param), pos_s, neg_s)
        ctx.save_for_backward(*tables)
        ctx.mark_non_differentiable(pos_s, neg_s)
        return loss, pos_s, neg_s

    @staticmethod
    def backward(ctx, grad_loss, _gp, _gn):
        cfg = ctx.cfg
        pos, corrupt, n_neg, batch_pos, loss_kind, param, pos_s, neg_s = ctx.args
        ph, pt, pr = pos
        tables = ctx.saved_tensors
        names = MODEL_TABLES[cfg.model]
        weights = dict(zip(names, tables))
        needs = dict(zip(names, ctx.needs_input_grad[8:]))
        n_pos = ph.numel()
        dev = ph.device
        gl = grad_loss.reshape(-1).contiguous().float()
        T = make_tables(weights, cfg.dim, cfg.l1, cfg.use_gumbel, cfg.item2ent)
        g = Grads()
        bufs = {}
        dense = cfg.grad_mode == "dense"
        g.mode = 1 if dense else 0
        shapes = {"ent": n_pos * (2 + n_neg), "rel": n_pos, "norm": n_pos}
        for name in names:
            buf = torch.zeros_like(weights[name]) if dense else \
                torch.empty((shapes[name], cfg.dim), dtype=torch.float32, device=dev)
            bufs[name] = buf
            setattr(g, name, buf.data_ptr())
        ent_ids = rel_ids = None
        if not dense:        # the COO index arrays come out of the same kernel pass
            ent_ids = torch.empty((1, n_pos * (2 + n_neg)), dtype=torch.int64, device=dev)
            rel_ids = torch.empty((1, n_pos), dtype=torch.int64, device=dev)
        lib = _lib.load()
        _lib.check(lib.kgrec_corrupt_loss_bwd(
            C.byref(T), cfg.model, _ptr(ph), _ptr(pt), _ptr(pr), _idx_bytes(ph, pt, pr), n_pos, _ptr(corrupt),
            n_neg, batch_pos, loss_kind, param, _ptr(pos_s), _ptr(neg_s), 1.0, _ptr(gl), C.byref(g),
            _ptr(ent_ids), _ptr(rel_ids), _stream()))
        count_launches(1)
        out = []
        for name in names:
            if not needs.get(name, False):
                out.append(None)
            elif dense:
                out.append(bufs[name])
            else:
                out.append(torch.sparse_coo_tensor(ent_ids if name == "ent" else rel_ids, bufs[name],
                                                   size=tuple(weights[name].shape), check_invariants=False))
        return (None,) * 8 + tuple(out)


def encode_corrupt(pos, neg):
    """(nh, nt, nr) -> the int32 corrupt format, for negatives that differ from their positive
    in exactly the head or the tail (what utils/data.py:12-56 produces).  Not validated here:
    use only on sampler output."""
    ph, pt, _ = pos
    nh, nt, _ = neg
    k = nh.numel() // ph.numel()
    head = nh.view(-1, k) != ph.view(-1, 1)
    c = torch.where(head, ~nh.view(-1, k).to(torch.int32), nt.view(-1, k).to(torch.int32))
    return c.contiguous().view(-1)


def corrupt_loss_step(cfg, weights, pos, corrupt, n_neg, batch_pos, loss_kind, param, status, grad_loss=1.0, reg=False):
    """Scores, per-batch losses and gradients of grad_loss * sum(loss) in one kernel
    (kgrec_corrupt_loss_step).  Returns (loss, pos_scores, neg_scores, {table: grad})."""
    names = MODEL_TABLES[cfg.model]
    T = make_tables(weights, cfg.dim, cfg.l1, cfg.use_gumbel, cfg.item2ent)
    ph, pt, pr = pos
    n_pos = ph.numel()
    dev = ph.device
    pos_s = torch.empty(n_pos, dtype=torch.float32, device=dev)
    neg_s = torch.empty(n_pos * n_neg, dtype=torch.float32, device=dev)
    loss = torch.empty((n_pos + batch_pos - 1) // batch_pos, dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws = torch.empty(lib.kgrec_corrupt_loss_step_workspace_bytes(C.byref(T), cfg.model, n_pos) // 4, dtype=torch.float32, device=dev)
    g = Grads()
    dense = cfg.grad_mode == "dense"
    g.mode = 1 if dense else 0
    shapes = {"ent": n_pos * (2 + n_neg), "rel": n_pos, "norm": n_pos}
    bufs = {}
    for name in names:
        bufs[name] = torch.zeros_like(weights[name]) if (dense or name not in shapes) else \
            torch.empty((shapes[name], cfg.dim), dtype=torch.float32, device=dev)      # TransR proj: dense accumulate
        setattr(g, name, bufs[name].data_ptr())
    # the COO index arrays of the slot gradients come out of the same kernel pass
    ent_ids = rel_ids = None
    if not dense:
        ent_ids = torch.empty((1, n_pos * (2 + n_neg)), dtype=torch.int64, device=dev)
        rel_ids = torch.empty((1, n_pos), dtype=torch.int64, device=dev)
    lib = _lib.load()
    _lib.check(lib.kgrec_corrupt_loss_step(
        C.byref(T), cfg.model, _ptr(ph), _ptr(pt), _ptr(pr), _idx_bytes(ph, pt, pr), n_pos, _ptr(corrupt), n_neg,
        batch_pos, loss_kind, float(param), float(grad_loss), 1 if reg else 0, _ptr(pos_s), _ptr(neg_s), _ptr(loss),
        C.byref(g), _ptr(ent_ids), _ptr(rel_ids), _ptr(ws), _ptr(status), _stream()))
    count_launches(2)
    grads = {}
    for name in names:
        if dense or name not in shapes:
            grads[name] = bufs[name]
        else:
            grads[name] = torch.sparse_coo_tensor(ent_ids if name == "ent" else rel_ids, bufs[name],
                                                  size=tuple(weights[name].shape), check_invariants=False)
    return loss, pos_s, neg_s, grads


def rank_loss_step(cfg, weights, pos, neg, n_neg, batch_pos, loss_kind, param, gumbel_u, status, grad_loss=1.0):
    """kgrec_rank_loss_step: scores, per-batch losses and the gradients of grad_loss * sum(loss) of the
    fused positive + K-negative ranking loss in one call (one kernel pass for TUP / KTUP pairs).
    Returns (loss, pos_scores, neg_scores, {table: grad})."""
    names = MODEL_TABLES[cfg.model]
    T = make_tables(weights, cfg.dim, cfg.l1, cfg.use_gumbel, cfg.item2ent)
    pa, pb, pc = pos
    na, nb, nc = neg
    n_pos = pa.numel()
    n = n_pos * (1 + n_neg)
    dev = pa.device
    pos_s = torch.empty(n_pos, dtype=torch.float32, device=dev)
    neg_s = torch.empty(n_pos * n_neg, dtype=torch.float32, device=dev)
    loss = torch.empty((n_pos + batch_pos - 1) // batch_pos, dtype=torch.float32, device=dev)
    ws = torch.empty(max(1, n_pos), dtype=torch.float32, device=dev)
    G, bufs = _alloc_grads(cfg.model, weights, n, cfg.grad_mode, cfg.dim)
    rec = cfg.model in (_lib.TUP, _lib.KTUP)
    sparse = cfg.grad_mode != "dense"
    ids = {}
    if rec and sparse:      # the COO index arrays come out of the same kernel pass
        for k in ("user", "item") + (("ent",) if cfg.model == _lib.KTUP else ()):
            ids[k] = torch.empty(n, dtype=torch.int64, device=dev)
    lib = _lib.load()
    _lib.check(lib.kgrec_rank_loss_step(
        C.byref(T), cfg.model, _ptr(pa), _ptr(pb), _ptr(pc), _ptr(na), _ptr(nb), _ptr(nc),
        _idx_bytes(pa, pb, pc, na, nb, nc), n_pos, n_neg, batch_pos, loss_kind, float(param), float(grad_loss),
        _ptr(gumbel_u), cfg.seed, _ptr(pos_s), _ptr(neg_s), _ptr(loss), C.byref(G),
        _ptr(ids.get("user")), _ptr(ids.get("item")), _ptr(ids.get("ent")), _ptr(ws), _ptr(status), _stream()))
    count_launches(2)
    idx = ids
    if sparse and not rec:
        pi = _slot_indices(cfg.model, pa, pb, pc, cfg.item2ent)
        ni = _slot_indices(cfg.model, na, nb, nc, cfg.item2ent)
        for k in pi:
            if k == "ent" and cfg.model != _lib.KTUP:
                idx[k] = torch.cat([pa, na, pb, nb]).long()
            else:
                idx[k] = torch.cat([pi[k], ni[k]])
    grads = _finish_grads(cfg.model, weights, bufs, idx, cfg.grad_mode, {k: True for k in names})
    return loss, pos_s, neg_s, grads


class GraphedCorruptStep:
    """The single-batch latency path: `corrupt_loss_step` (forward + ranking loss [+ regularisers] + backward, two
    kernels) captured ONCE in a CUDA graph over static buffers and replayed per batch.

    A reference-sized batch (1024 positives + 10 negatives each) is ~5 us of GPU work; called through the module
    API it costs ~80 us of host time (index conversion, ten allocations, the ctypes struct, two launches, two sparse
    tensor wrappers).  Here the per-step host work is writing the batch's ids into `self.h / .t / .r / .corrupt`
    (int32, device; e.g. the target of the H2D copy) and one `cudaGraphLaunch`.  Outputs live in `self.loss`
    (per loss batch), `self.pos_scores`, `self.neg_scores` and `self.grads` ({table: sparse COO slots | dense}),
    overwritten by every replay.  Shapes, loss parameters and the tables' storage are fixed at capture."""

    def __init__(self, model, n_pos, n_neg, margin=1.0, loss="margin", batch_pos=None, grad_loss=1.0, reg=False,
                 kg_branch_model=None):
        dev = model._require_cuda()
        kmodel = model.MODEL if kg_branch_model is None else kg_branch_model
        names = MODEL_TABLES[kmodel]
        w = model._weights()
        self.model = model
        self.weights = {k: w[k] for k in names}
        self.h, self.t, self.r = (torch.zeros(n_pos, dtype=torch.int32, device=dev) for _ in range(3))
        self.corrupt = torch.zeros(n_pos * n_neg, dtype=torch.int32, device=dev)
        kind = {"margin": _lib.LOSS_MARGIN, "bpr": _lib.LOSS_BPR}[loss]
        cfg = _Ctx(kmodel, model.embedding_size, model.L1_flag, False, None, model.grad_mode, 0)
        status = model._status_buf(dev)

        def run():
            return corrupt_loss_step(cfg, self.weights, (self.h, self.t, self.r), self.corrupt, n_neg, batch_pos or n_pos,
                                     kind, margin, status, grad_loss, reg)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            run()                                          # warm-up outside the graph (lazy module loading)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.pos_scores, self.neg_scores, self.grads = run()
        self.launches_per_replay = 2

    def replay(self):
        self.graph.replay()
        count_launches(self.launches_per_replay)
        return self.loss
