"""Sparse-row training step: fused forward + loss + backward into persistent accumulators,
then an optimizer update of exactly the rows the batch touched.

Replaces, for the KG models, the reference's `trainer.optimizer_zero_grad(); losses.backward();
clip_grad_norm; trainer.optimizer_step()` sequence (knowledge_representation.py:187-216,
utils/trainer.py:63-81) whose cost is O(table) per step with one whose cost is O(batch).
Update rules are torch.optim's SGD / Adagrad / Adam formulas; Adam is applied to the touched
rows only ("lazy"), which differs from dense Adam on untouched rows.
"""
import ctypes as C

import torch

from . import _lib
from . import functional as KF

_KINDS = {"SGD": 0, "Adagrad": 1, "Adam": 2}


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
        self.flags = {k: torch.zeros(w[k].shape[0], dtype=torch.int32, device=dev) for k in self.names}
        self.s1 = {k: torch.zeros_like(w[k]) for k in self.names} if self.kind else {k: None for k in self.names}
        self.s2 = {k: torch.zeros_like(w[k]) for k in self.names} if self.kind == 2 else {k: None for k in self.names}
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)

    def step_corrupt(self, pos, corrupt, margin=1.0, loss="margin", batch_pos=None, reg=False):
        """One training step on positives (h, t, r) and group-compact negatives; returns the
        per-batch losses (device tensor; nothing synchronises).  reg=True adds the KG drivers'
        normLoss / orthogonalLoss regularisers inside the same kernel (kgrec_corrupt_loss_step)."""
        m = self.model
        if m.MODEL not in (_lib.TRANSE, _lib.TRANSH):
            raise NotImplementedError("SparseRowOptimizer.step_corrupt is built for TransE / TransH")
        dev = m._require_cuda()
        pos = tuple(KF.as_index(x, dev) for x in pos)
        corrupt = corrupt.to(dev, torch.int32, non_blocking=True).contiguous().view(-1)
        n_pos = pos[0].numel()
        n_neg = corrupt.numel() // n_pos
        w = m._weights()
        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ptr = KF._ptr
        # forward + loss + backward, gradients atomically accumulated into the persistent buffers
        T = KF.make_tables({k: w[k] for k in self.names}, m.embedding_size, m.L1_flag)
        g = _lib.Grads()
        g.mode = 1
        for k in self.names:
            setattr(g, k, self.acc[k].data_ptr())
        pos_s = torch.empty(n_pos, dtype=torch.float32, device=dev)
        neg_s = torch.empty(n_pos * n_neg, dtype=torch.float32, device=dev)
        bp = batch_pos or n_pos
        out = torch.empty((n_pos + bp - 1) // bp, dtype=torch.float32, device=dev)
        ws = torch.empty(max(1, n_pos), dtype=torch.float32, device=dev)
        kind = {"margin": _lib.LOSS_MARGIN, "bpr": _lib.LOSS_BPR}[loss]
        # the rows this batch touched come out of the same kernel pass (slot row ids)
        ent_ids = torch.empty(n_pos * (2 + n_neg), dtype=torch.int64, device=dev)
        rel_ids = torch.empty(n_pos, dtype=torch.int64, device=dev)
        _lib.check(lib.kgrec_corrupt_loss_step(
            C.byref(T), m.MODEL, ptr(pos[0]), ptr(pos[1]), ptr(pos[2]), pos[0].element_size(), n_pos, ptr(corrupt),
            n_neg, bp, kind, float(margin), 1.0, 1 if reg else 0, ptr(pos_s), ptr(neg_s), ptr(out), C.byref(g), ptr(ent_ids), ptr(rel_ids),
            ptr(ws), ptr(m._status_buf(dev)), stream))
        KF.count_launches(2)
        ids = {"ent": ent_ids, "rel": rel_ids, "norm": rel_ids}
        self.t += 1
        use_clip = self.clip is not None
        if use_clip:
            self.sqnorm.zero_()
            for k in self.names:
                _lib.check(lib.kgrec_rows_sqnorm(ptr(self.acc[k]), ptr(self.flags[k]), ptr(ids[k]), ids[k].element_size(),
                                                 ids[k].numel(), w[k].shape[0], w[k].shape[1], ptr(self.sqnorm), stream))
            KF.count_launches(len(self.names))
        for k in self.names:
            _lib.check(lib.kgrec_rows_step(
                ptr(w[k].data), ptr(self.acc[k]), ptr(self.s1[k]), ptr(self.s2[k]), ptr(self.flags[k]), ptr(ids[k]),
                ids[k].element_size(), ids[k].numel(), w[k].shape[0], w[k].shape[1], self.kind, self.lr, self.eps,
                self.betas[0], self.betas[1], self.t, self.wd, ptr(self.sqnorm) if use_clip else None,
                float(self.clip or 0.0), 1 if use_clip else 0, stream))
        KF.count_launches(2 * len(self.names))
        return out
