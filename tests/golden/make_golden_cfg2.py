"""Golden vectors at a BASELINE config shape: configs[1] (TransE / TransH, d=100, |E|=100k, |R|=500, one batch of
1024 positives with 10 corrupted negatives each), recorded from the UNMODIFIED reference classes in /root/reference
with the call pattern the reference needs for K negatives per positive -- model(pos.repeat_interleave(10)) vs
model(neg) + marginLoss + backward (SURVEY 8d cfg#2 note).

    python tests/golden/make_golden_cfg2.py          (build container only; outputs are committed)

The 40 MB tables are not stored: the drop-in constructors consume torch's generator exactly as the reference
constructors do (tests/test_drivers.py::test_dropin_constructors_start_from_the_reference_tables), so the test rebuilds
them from the seed and checks a few recorded rows first.  Stored: ids, scores, loss, the dense relation(-side) gradients and
the entity-gradient rows of 512 touched entities.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("KGREC_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
if not hasattr(np, "asfarray"):
    np.asfarray = lambda a, dtype=float: np.asarray(a, dtype=dtype)  # noqa: E731
sys.path.insert(0, REF)
from jTransUP.models.transE import TransEModel      # noqa: E402
from jTransUP.models.transH import TransHModel      # noqa: E402
from jTransUP.utils import loss as ref_loss         # noqa: E402

torch.set_num_threads(8)
D, E, R, B, K, SEED = 100, 100_000, 500, 1024, 10, 7

for name, cls, l1 in (("transe_l2", TransEModel, False), ("transe_l1", TransEModel, True), ("transh_l2", TransHModel, False)):
    torch.manual_seed(SEED)
    m = cls(l1, D, E, R)
    g = torch.Generator().manual_seed(SEED + 1)
    ph, pt = torch.randint(0, E, (B,), generator=g), torch.randint(0, E, (B,), generator=g)
    pr = torch.randint(0, R, (B,), generator=g)
    c = torch.randint(0, E, (B * K,), generator=g)
    head = torch.rand(B * K, generator=g) < 0.5
    rep = lambda x: x.repeat_interleave(K)          # noqa: E731
    nh, nt, nr = torch.where(head, c, rep(ph)), torch.where(head, rep(pt), c), rep(pr)
    pos = m(rep(ph), rep(pt), rep(pr))              # every positive scored K times, paired element-wise with its negatives
    neg = m(nh, nt, nr)
    loss = ref_loss.marginLoss()(pos, neg, 1.0)
    loss.backward()
    ent_grad = m.ent_embeddings.weight.grad
    touched = torch.unique(torch.cat([ph, pt, c]))
    sample = touched[torch.randperm(touched.numel(), generator=g)[:512]]
    out = {
        "seed": np.int64(SEED), "ph": ph.numpy(), "pt": pt.numpy(), "pr": pr.numpy(),
        "corrupt": torch.where(head, ~c, c).to(torch.int32).numpy(),      # group-compact format of the same negatives
        "pos_scores": pos.detach().view(B, K)[:, 0].numpy(), "neg_scores": neg.detach().numpy(), "loss": np.float32(loss.item()),
        "ent_rows_check": m.ent_embeddings.weight.detach()[:4].numpy(), "rel_rows_check": m.rel_embeddings.weight.detach()[-4:].numpy(),
        "ent_grad_ids": sample.numpy(), "ent_grad_rows": ent_grad[sample].numpy(),
        "ent_grad_sqnorm": np.float64((ent_grad.double() ** 2).sum().item()),
        "rel_grad": m.rel_embeddings.weight.grad.numpy(),
    }
    if cls is TransHModel:
        out["norm_grad"] = m.norm_embeddings.weight.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "cfg2_%s.npz" % name), **out)
    print(name, float(loss), "bytes", os.path.getsize(os.path.join(OUT, "cfg2_%s.npz" % name)))
