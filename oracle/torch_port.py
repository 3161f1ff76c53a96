"""Torch-CPU port of the reference's op sequence for the hot path (TEST / BASELINE
INFRASTRUCTURE ONLY -- see oracle/kg_oracle.py's header; never on the product path).

Where kg_oracle.py restates the arithmetic in numpy, this file restates the *same chain of
stock torch ops the reference executes* (nn.Embedding gathers, pointwise ops, sum, the
[B, N, d] broadcast in evaluate*, autograd with dense [rows, d] gradients), so that timing it
on the host cores is the closest thing to "the reference's own PyTorch-CPU path" that can
travel to the GPU box (the reference checkout cannot).  It is validated against the same
golden vectors in tests/test_oracle_golden.py.  Citations: transE.py:51-105,
transH.py:58-121, utils/misc.py:18-19, utils/loss.py:8-16.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _proj(x, w):                                   # utils/misc.py:18-19
    return x - torch.sum(x * w, dim=x.dim() - 1, keepdim=True) * w


class TransPort(nn.Module):
    """TransE (norm=False) / TransH (norm=True) with the reference's layer structure."""

    def __init__(self, l1, d, ent_total, rel_total, with_norm):
        super().__init__()
        self.l1, self.d, self.ent_total, self.with_norm = l1, d, ent_total, with_norm
        self.ent_embeddings = nn.Embedding(ent_total, d)
        self.rel_embeddings = nn.Embedding(rel_total, d)
        if with_norm:
            self.norm_embeddings = nn.Embedding(rel_total, d)
        with torch.no_grad():
            for emb in self.children():
                nn.init.xavier_uniform_(emb.weight)
                emb.weight.copy_(F.normalize(emb.weight, p=2, dim=1))

    def _dist(self, e, dim):
        return torch.sum(torch.abs(e), dim) if self.l1 else torch.sum(e ** 2, dim)

    def forward(self, h, t, r):                    # transE.py:51-63 / transH.py:58-71
        h_e, t_e, r_e = self.ent_embeddings(h), self.ent_embeddings(t), self.rel_embeddings(r)
        if self.with_norm:
            w = self.norm_embeddings(r)
            h_e, t_e = _proj(h_e, w), _proj(t_e, w)
        return self._dist(h_e + r_e - t_e, 1)

    def evaluate_side(self, q, r, head):           # transE.py:65-105 / transH.py:73-121
        b = len(q)
        q_e, r_e = self.ent_embeddings(q), self.rel_embeddings(r)
        ent = self.ent_embeddings.weight.expand(b, self.ent_total, self.d)
        if self.with_norm:
            w = self.norm_embeddings(r)
            q_e = _proj(q_e, w)
            ent = _proj(ent, w.expand(self.ent_total, b, self.d).permute(1, 0, 2))
        c = q_e - r_e if head else q_e + r_e
        c = c.expand(self.ent_total, b, self.d).permute(1, 0, 2)
        return self._dist(c - ent, 2)


def margin_loss(pos, neg, margin):                 # utils/loss.py:12-16
    return torch.sum(torch.max(pos - neg + margin, torch.zeros_like(pos)))


def train_step(model, pos, neg, margin=1.0):
    """forward(pos) + forward(neg) + marginLoss + backward with dense gradients, as
    knowledge_representation.py:189-207 does per step.  K negatives per positive are fed the
    way the reference would see them: the positives repeated K times."""
    ph, pt, pr = pos
    nh, nt, nr = neg
    k = nh.numel() // ph.numel()
    model.zero_grad(set_to_none=True)
    ps = model(ph, pt, pr)
    ns = model(nh, nt, nr)
    loss = margin_loss(ps.repeat_interleave(k) if k > 1 else ps, ns, margin)
    loss.backward()
    return loss
