"""Sparse-row training step: fused forward + loss + backward into persistent accumulators, then
a clip + optimizer update of exactly the rows the batch touched.

Replaces the reference's `trainer.optimizer_zero_grad(); losses.backward(); clip_grad_norm;
trainer.optimizer_step()` sequence (knowledge_representation.py:187-216, item_recommendation.py:
168-192, knowledgable_recommendation.py:320-402, utils/trainer.py:63-81), whose cost is O(table)
per step, with one whose cost is O(batch):

    step kernel (dense-accumulate gradients)      kgrec_corrupt_loss_step / kgrec_rank_loss_step
    [regularisers of the driver]                  fused (KG side) / kgrec_reg_* (rec side)
    epoch marks from the batch's id arrays        kgrec_rows_mark        1 launch
    clip_grad_norm's total norm                   kgrec_rows_sqnorm      1 launch, all tables
    SGD / Adagrad / Adam on the marked rows       kgrec_rows_update      1 launch, all tables

Update rules are torch.optim's SGD / Adagrad / Adam formulas; Adam is applied to the touched rows
only ("lazy"), which differs from dense Adam on untouched rows.  A table takes
part in a step when the step's loss reaches it -- as in the reference, where parameters whose
``.grad`` is None are skipped (KTUP: the rec branch moves user / item / aligned entities / pref /
pref_norm / rel / norm, the KG branch ent / rel / norm).
"""
import ctypes as C

import torch

from . import _lib
from . import functional as KF

_KINDS = {"SGD": 0, "Adagrad": 1, "Adam": 2}
_ATTR = {"ent": "ent_embeddings", "rel": "rel_embeddings", "norm": "norm_embeddings", "proj": "proj_embeddings",
         "user": "user_embeddings", "item": "item_embeddings", "pref": "pref_embeddings",
         "pref_norm": "pref_norm_embeddings"}
# tables whose rows are gathered by id (the others are small and every row takes part)
_GATHERED = ("ent", "rel", "norm", "user", "item")


class SparseRowOptimizer:
    def __init__(self, model, optimizer_type="Adagrad", lr=0.01, l2_lambda=0.0, clip=None,
                 eps=None, betas=(0.9, 0.999)):
        if optimizer_type not in _KINDS:
            raise ValueError("optimizer_type must be one of %s" % sorted(_KINDS))
        self.model, self.kind, self.lr, self.wd, self.clip = model, _KINDS[optimizer_type], lr, l2_lambda, clip
        self.eps = eps if eps is not None else (1e-10 if self.kind == 1 else 1e-8)
        self.betas = betas
        self.t = 0
        dev = model._require_cuda()
        self.names = KF.MODEL_TABLES[model.MODEL]
        w = model._weights()
        self.acc = {k: torch.zeros_like(w[k]) for k in self.names}
        ktup = model.MODEL == _lib.KTUP
        self.marks = {k: torch.zeros(w[k].shape[0], dtype=torch.int32, device=dev)
                      for k in self.names if k in _GATHERED and not (ktup and k in ("rel", "norm"))}
        self.s1 = {k: torch.zeros_like(w[k]) for k in self.names} if self.kind else {k: None for k in self.names}
        self.s2 = {k: torch.zeros_like(w[k]) for k in self.names} if self.kind == 2 else {k: None for k in self.names}
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.reg_loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._rows_ws = None          # work buffers of the row-factored soft rec step, allocated on first use

    # -- shared tail of every step: marks -> total norm -> update --------------------------------------
    def _seg(self, ids, table, compact=False, remap=None):
        m = self.marks[table]
        return _lib.MarkSeg(ids=ids.data_ptr(), n=ids.numel(), idx_bytes=ids.element_size(), compact=int(compact),
                            remap=remap.data_ptr() if remap is not None else None,
                            n_remap=remap.numel() if remap is not None else 0,
                            marks=m.data_ptr(), rows=m.numel())

    def _mark(self, segs):
        """Start a step: bump the epoch and mark the rows the step's id arrays name."""
        m = self.model
        self.t += 1
        seg_arr = (_lib.MarkSeg * len(segs))(*segs)
        _lib.check(_lib.load().kgrec_rows_mark(seg_arr, len(segs), self.t, KF._ptr(m._status_buf(m.device)), KF._stream()))
        KF.count_launches(1)

    def _update(self, tables):
        """Finish a step: total norm (clip) and the optimizer update of the marked rows of `tables`."""
        m = self.model
        lib = _lib.load()
        stream = KF._stream()
        w = m._weights()
        entries = []
        for k in tables:
            mk = self.marks.get(k)
            entries.append(_lib.OptTable(
                table=w[k].data.data_ptr(), acc=self.acc[k].data_ptr(),
                state1=self.s1[k].data_ptr() if self.s1[k] is not None else None,
                state2=self.s2[k].data_ptr() if self.s2[k] is not None else None,
                marks=mk.data_ptr() if mk is not None else None, rows=w[k].shape[0], dim=w[k].shape[1], keep_acc=0))
        tab_arr = (_lib.OptTable * len(entries))(*entries)
        use_clip = self.clip is not None
        if use_clip:
            self.sqnorm.zero_()
            _lib.check(lib.kgrec_rows_sqnorm(tab_arr, len(entries), self.t, KF._ptr(self.sqnorm), stream))
        _lib.check(lib.kgrec_rows_update(tab_arr, len(entries), self.t, self.kind, self.lr, self.eps, self.betas[0],
                                         self.betas[1], self.t, self.wd, KF._ptr(self.sqnorm) if use_clip else None,
                                         float(self.clip or 0.0), stream))
        KF.count_launches(1 + int(use_clip))

    def _apply(self, segs, tables):
        """segs: MarkSeg list of this step's id arrays; tables: names of the tables the step's loss reaches."""
        self._mark(segs)
        self._update(tables)

    def _grads(self, names):
        g = _lib.Grads()
        g.mode = 1
        for k in names:
            setattr(g, k, self.acc[k].data_ptr())
        return g

    # -- KG models: TransE / TransH / TransR, and the KG branch of KTUP ---------------------------------
    def step_corrupt(self, pos, corrupt, margin=1.0, loss="margin", batch_pos=None, reg=False, grad_loss=1.0):
        """One training step on positives (h, t, r) and group-compact negatives; returns the
        per-batch losses (device tensor; nothing synchronises).  reg=True adds the KG drivers'
        normLoss / orthogonalLoss regularisers inside the same kernel (kgrec_corrupt_loss_step).
        KTUP: the joint model's KG branch (TransH on ent / rel / norm), grad_loss = kg_lambda."""
        m = self.model
        if m.MODEL not in (_lib.TRANSE, _lib.TRANSH, _lib.TRANSR, _lib.KTUP):
            raise NotImplementedError("step_corrupt: KG models and the KG branch of KTUP")
        kmodel = _lib.TRANSH if m.MODEL == _lib.KTUP else m.MODEL
        names = KF.MODEL_TABLES[kmodel]
        dev = m._require_cuda()
        pos = tuple(KF.as_index(x, dev) for x in pos)
        idx_bytes = KF._idx_bytes(*pos)
        corrupt = corrupt.to(dev, torch.int32, non_blocking=True).contiguous().view(-1)
        n_pos = pos[0].numel()
        bp = batch_pos or max(1, n_pos)
        out = torch.zeros((n_pos + bp - 1) // bp if n_pos else 0, dtype=torch.float32, device=dev)
        if n_pos == 0:
            return out
        if corrupt.numel() % n_pos:
            raise ValueError("corrupt ids must be a whole multiple of the positives")
        n_neg = corrupt.numel() // n_pos
        w = m._weights()
        lib = _lib.load()
        ptr = KF._ptr
        T = KF.make_tables({k: w[k] for k in names}, m.embedding_size, m.L1_flag)
        g = self._grads(names)
        pos_s = torch.empty(n_pos, dtype=torch.float32, device=dev)
        neg_s = torch.empty(n_pos * n_neg, dtype=torch.float32, device=dev)
        ws = torch.empty(lib.kgrec_corrupt_loss_step_workspace_bytes(C.byref(T), kmodel, n_pos) // 4, dtype=torch.float32, device=dev)
        kind = {"margin": _lib.LOSS_MARGIN, "bpr": _lib.LOSS_BPR}[loss]
        _lib.check(lib.kgrec_corrupt_loss_step(
            C.byref(T), kmodel, ptr(pos[0]), ptr(pos[1]), ptr(pos[2]), idx_bytes, n_pos, ptr(corrupt),
            n_neg, bp, kind, float(margin), float(grad_loss), 1 if reg else 0, ptr(pos_s), ptr(neg_s), ptr(out),
            C.byref(g), None, None, ptr(ws), ptr(m._status_buf(dev)), KF._stream()))
        KF.count_launches(2)
        segs = [self._seg(pos[0], "ent"), self._seg(pos[1], "ent"), self._seg(corrupt, "ent", compact=True)]
        segs += [self._seg(pos[2], k) for k in ("rel", "norm") if k in names and k in self.marks]   # KTUP: small tables, all rows
        self._apply(segs, names)
        return out

    # -- recommendation models: TUP, and the rec branch of KTUP -----------------------------------------------
    def step_pairs(self, pos, neg, target=-1.0, loss="bpr", batch_pos=None, gumbel_u=None, reg=False):
        """One training step on (u, i) positives and (u repeated, ni) negatives: forward + ranking
        loss + backward in the tile kernel (kgrec_rank_loss_step), then clip + update.  reg=True adds
        the driver's regularisers: TUP (item_recommendation.py:177-180) orthogonalLoss(pref, pref_norm) +
        normLoss(user rows) + normLoss(item rows of cat[pos, neg]) + normLoss(pref); KTUP rec branch
        (knowledgable_recommendation.py:343-344) orthogonalLoss(pref, pref_norm).
        Returns (loss per batch, regulariser value) as device tensors."""
        m = self.model
        if m.MODEL not in (_lib.TUP, _lib.KTUP):
            raise NotImplementedError("step_pairs: TUP / KTUP")
        dev = m._require_cuda()
        pu, pi = (KF.as_index(x, dev) for x in pos)
        nu, ni = (KF.as_index(x, dev) for x in neg)
        idx_bytes = KF._idx_bytes(pu, pi, nu, ni)
        n_pos = pu.numel()
        bp = batch_pos or max(1, n_pos)
        out = torch.zeros((n_pos + bp - 1) // bp if n_pos else 0, dtype=torch.float32, device=dev)
        self.reg_loss.zero_()
        if n_pos == 0:
            return out, self.reg_loss
        if ni.numel() % n_pos:
            raise ValueError("negatives must be a whole multiple of the positives")
        n_neg = ni.numel() // n_pos
        names = self.names
        w = m._weights()
        lib = _lib.load()
        ptr = KF._ptr
        stream = KF._stream()
        T = KF.make_tables({k: w[k] for k in names}, m.embedding_size, m.L1_flag, m.use_st_gumbel, m._item2ent)
        ktup = m.MODEL == _lib.KTUP
        g = self._grads([k for k in names if not (ktup and k in ("rel", "norm"))])   # KTUP: rel / norm share pref / pref_norm's
        pos_s = torch.empty(n_pos, dtype=torch.float32, device=dev)
        neg_s = torch.empty(n_pos * n_neg, dtype=torch.float32, device=dev)
        ws = torch.empty(max(1, n_pos), dtype=torch.float32, device=dev)
        kind = {"margin": _lib.LOSS_MARGIN, "bpr": _lib.LOSS_BPR}[loss]
        if gumbel_u is not None:
            gumbel_u = gumbel_u.to(dev, torch.float32).contiguous()
        seed = m._next_seed() if (m.use_st_gumbel and gumbel_u is None) else 0
        segs = [self._seg(pu, "user"), self._seg(pi, "item"), self._seg(ni, "item")]
        if ktup:
            segs += [self._seg(pi, "ent", remap=m._item2ent), self._seg(ni, "ent", remap=m._item2ent)]
        self._mark(segs)
        rows_path = self._use_rows_path(n_pos, n_neg, nu, pu)
        if rows_path:
            # soft preferences: the [P x d] contractions once per distinct row of the step (csrc/train_rec_rows.cu)
            if self._rows_ws is None:
                n_fl = lib.kgrec_rec_rows_workspace_floats(w["user"].shape[0], w["item"].shape[0], m.embedding_size,
                                                           w["pref"].shape[0], 1 if ktup else 0)
                self._rows_ws = torch.empty(int(n_fl), dtype=torch.float32, device=dev)
                first = 1
            else:
                first = 0
            _lib.check(lib.kgrec_rec_rows_step(
                C.byref(T), m.MODEL, ptr(pu), ptr(pi), ptr(ni), idx_bytes, n_pos, n_neg, bp, kind, float(target), 1.0,
                ptr(self.marks["user"]), ptr(self.marks["item"]), self.t, ptr(self._rows_ws), first, C.byref(g),
                ptr(pos_s), ptr(neg_s), ptr(out), ptr(ws),
                ptr(self.reg_loss) if (reg and not ktup) else None,
                ptr(gumbel_u), seed, ptr(m._status_buf(dev)), stream))
            KF.count_launches(8)
        else:
            _lib.check(lib.kgrec_rank_loss_step(
                C.byref(T), m.MODEL, ptr(pu), ptr(pi), None, ptr(nu), ptr(ni), None, idx_bytes, n_pos, n_neg, bp, kind,
                float(target), 1.0, ptr(gumbel_u), seed, ptr(pos_s), ptr(neg_s), ptr(out), C.byref(g),
                None, None, None, ptr(ws), ptr(m._status_buf(dev)), stream))
            KF.count_launches(2)
        if ktup:      # (pref + rel) and (pref_norm + norm) enter the score as sums (jTransUP.py:253-258): equal gradients
            self.acc["rel"].copy_(self.acc["pref"])
            self.acc["norm"].copy_(self.acc["pref_norm"])
        if reg:
            d = m.embedding_size
            pw, nw = w["pref"], w["pref_norm"]
            _lib.check(lib.kgrec_reg_orth_tables(ptr(pw), ptr(nw), pw.shape[0], d, 1.0, ptr(self.reg_loss),
                                                 ptr(self.acc["pref"]), ptr(self.acc["pref_norm"]), stream))
            KF.count_launches(1)
            if not ktup:
                status = ptr(m._status_buf(dev))
                fused = rows_path                                  # rows path: normLoss of the gathered rows is fused into its pair kernel
                for tab, ids in (() if fused else (("user", pu), ("item", pi), ("item", ni))):
                    _lib.check(lib.kgrec_reg_norm_rows(ptr(w[tab]), w[tab].shape[0], d, ptr(ids), ids.element_size(),
                                                       ids.numel(), 1.0, ptr(self.reg_loss), ptr(self.acc[tab]), status, stream))
                _lib.check(lib.kgrec_reg_norm_rows(ptr(pw), pw.shape[0], d, None, 8, pw.shape[0], 1.0, ptr(self.reg_loss),
                                                   ptr(self.acc["pref"]), None, stream))
                KF.count_launches(4)
        self._update(names)
        return out, self.reg_loss

    def _use_rows_path(self, n_pos, n_neg, nu, pu):
        """Row-factored soft step (train_rec_rows.cu) when the step re-uses rows: its per-row kernels cost about one
        pair-kernel pair per DISTINCT row, its pair kernel about a third of one.  KGREC_REC_ROWS=0 | force overrides."""
        import os
        m = self.model
        env = os.environ.get("KGREC_REC_ROWS", "")
        d, P = m.embedding_size, m.pref_embeddings.weight.shape[0]
        ok = d % 4 == 0 and d <= 128 and P <= 32 and 1 <= n_neg <= 31 and not (m.use_st_gumbel and m.L1_flag)
        if not ok or env == "0":
            return False
        # the row path scores negative k of positive j as (pu[j], ni[j, k]): the (u repeated, ni) contract of this
        # method (what getNegRatings produces); nu itself is not read
        if env == "force":
            return True
        pairs = n_pos * (1 + n_neg)
        rows = min(m.user_embeddings.weight.shape[0], n_pos) + min(m.item_embeddings.weight.shape[0], pairs)
        return rows <= 0.35 * pairs          # measured: ~3.2 ns per row vs 2.1 (pair kernel) - 0.7 (row path) ns per pair
