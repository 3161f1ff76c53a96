"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; it cannot travel to the
GPU box, which is why the outputs are committed):

    python tests/golden/make_golden.py

It imports the reference model classes and helper functions from
/root/reference (nothing is copied), instantiates them on CPU with fixed seeds
at small sizes, and records inputs, weights, forward scores, full-catalog score
matrices, loss values, dense parameter gradients and ranking results.  The
Gumbel uniform draw is captured by re-seeding torch's global generator exactly
as SURVEY.md section 7.3-2 describes (the reference draws
``logits.data.new(*logits.size()).uniform_()`` and nothing else consumes the RNG).

Shims (on the fly, no file of the reference is edited): ``numpy.asfarray``
(removed in numpy 2, used at utils/evaluation.py:69).
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
from jTransUP.models.transR import TransRModel      # noqa: E402
from jTransUP.models.transUP import TransUPModel    # noqa: E402
from jTransUP.models.jTransUP import jTransUPModel  # noqa: E402
from jTransUP.utils import loss as ref_loss         # noqa: E402
from jTransUP.utils.misc import getKGPerformance, getRecPerformance  # noqa: E402

torch.set_num_threads(1)
D, E, R, U, I, P, B = 20, 37, 5, 11, 23, 4, 9
LT = torch.LongTensor


def npy(t):
    return t.detach().cpu().numpy().copy()


def grads_of(model):
    return {"grad_" + k.replace(".weight", ""): npy(p.grad) if p.grad is not None
            else np.zeros(tuple(p.shape), np.float32) for k, p in model.named_parameters()}


def weights_of(model):
    return {"w_" + k.replace(".weight", ""): npy(v) for k, v in model.state_dict().items()}


def kg_case(name, cls, l1, seed):
    torch.manual_seed(seed)
    m = cls(L1_flag=l1, embedding_size=D, ent_total=E, rel_total=R)
    rng = np.random.RandomState(seed)
    ph, pt, nh, nt = (rng.randint(0, E, B) for _ in range(4))
    pr = rng.randint(0, R, B)
    nr = pr.copy()
    out = dict(weights_of(m), l1=l1, ph=ph, pt=pt, pr=pr, nh=nh, nt=nt, nr=nr, margin=1.0)
    pos = m(LT(ph), LT(pt), LT(pr))
    neg = m(LT(nh), LT(nt), LT(nr))
    loss = ref_loss.marginLoss()(pos, neg, 1.0)
    loss.backward()
    out.update(pos=npy(pos), neg=npy(neg), loss=npy(loss), **grads_of(m))
    # plain sum-of-scores gradient (upstream of ones) for the score backward alone
    m.zero_grad()
    m(LT(ph), LT(pt), LT(pr)).sum().backward()
    out.update({k.replace("grad_", "gsum_"): v for k, v in grads_of(m).items()})
    q, qr = rng.randint(0, E, 4), rng.randint(0, R, 4)
    out.update(q=q, qr=qr, eval_head=npy(m.evaluateHead(LT(q), LT(qr))),
               eval_tail=npy(m.evaluateTail(LT(q), LT(qr))))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def gumbel_draw(seed, shape):
    torch.manual_seed(seed)
    return torch.empty(*shape).uniform_()


def tup_case(name, l1, gumbel, seed):
    torch.manual_seed(seed)
    m = TransUPModel(L1_flag=l1, embedding_size=D, user_total=U, item_total=I,
                     preference_total=P, use_st_gumbel=gumbel)
    rng = np.random.RandomState(seed)
    u, pi, ni = rng.randint(0, U, B), rng.randint(0, I, B), rng.randint(0, I, B)
    out = dict(weights_of(m), l1=l1, gumbel=gumbel, u=u, pi=pi, ni=ni, target=-1.0)
    s1, s2, s3 = seed + 100, seed + 200, seed + 300
    if gumbel:
        out.update(noise_pos=npy(gumbel_draw(s1, (B, P))), noise_neg=npy(gumbel_draw(s2, (B, P))),
                   noise_eval=npy(gumbel_draw(s3, (3, I, P))))
    torch.manual_seed(s1)
    pos = m(LT(u), LT(pi))
    torch.manual_seed(s2)
    neg = m(LT(u), LT(ni))
    loss = ref_loss.bprLoss(pos, neg, target=-1.0)
    loss.backward()
    out.update(pos=npy(pos), neg=npy(neg), loss=npy(loss), **grads_of(m))
    m.zero_grad()
    torch.manual_seed(s1)
    m(LT(u), LT(pi)).sum().backward()
    out.update({k.replace("grad_", "gsum_"): v for k, v in grads_of(m).items()})
    qu = rng.randint(0, U, 3)
    torch.manual_seed(s3)
    out.update(qu=qu, eval=npy(m.evaluate(LT(qu))))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def ktup_case(name, l1, gumbel, seed):
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    # joint vocabulary as load_kg_rating_data.rebuildEntityItemVocab builds it:
    # i_map: item -> joint index ; new_map: joint index -> (entity or -1, item or -1)
    aligned = rng.rand(I) < 0.7
    ents = rng.permutation(E)[:I]
    i_map = {i: i for i in range(I)}
    new_map = {i: ((int(ents[i]) if aligned[i] else -1), i) for i in range(I)}
    m = jTransUPModel(L1_flag=l1, embedding_size=D, user_total=U, item_total=I, entity_total=E,
                      relation_total=R, i_map=i_map, new_map=new_map, isShare=False,
                      use_st_gumbel=gumbel)
    item2ent = np.array([new_map[i_map[i]][0] if new_map[i_map[i]][0] != -1 else E for i in range(I)])
    u, pi, ni = rng.randint(0, U, B), rng.randint(0, I, B), rng.randint(0, I, B)
    out = dict(weights_of(m), l1=l1, gumbel=gumbel, u=u, pi=pi, ni=ni, target=-1.0,
               item2ent=item2ent)
    s1, s2, s3 = seed + 100, seed + 200, seed + 300
    if gumbel:
        out.update(noise_pos=npy(gumbel_draw(s1, (B, R))), noise_neg=npy(gumbel_draw(s2, (B, R))),
                   noise_eval=npy(gumbel_draw(s3, (3, I, R))))
    # the reference iterates a LongTensor and indexes a dict with 0-d tensors
    # (KeyError on modern torch); feed python ints through a list-like
    # wrapper instead of editing the reference: paddingItems only iterates.
    orig_pad = m.paddingItems
    m.paddingItems = lambda ids, pad: orig_pad([int(x) for x in ids], pad)
    torch.manual_seed(s1)
    pos = m((LT(u), LT(pi)), None, is_rec=True)
    torch.manual_seed(s2)
    neg = m((LT(u), LT(ni)), None, is_rec=True)
    loss = ref_loss.bprLoss(pos, neg, target=-1.0)
    loss.backward()
    out.update(pos=npy(pos), neg=npy(neg), loss=npy(loss), **grads_of(m))
    # KG branch
    m.zero_grad()
    ph, pt, nh, nt = (rng.randint(0, E, B) for _ in range(4))
    pr = rng.randint(0, R, B)
    kpos = m(None, (LT(ph), LT(pt), LT(pr)), is_rec=False)
    kneg = m(None, (LT(nh), LT(nt), LT(pr)), is_rec=False)
    kloss = ref_loss.marginLoss()(kpos, kneg, 1.0)
    kloss.backward()
    out.update(ph=ph, pt=pt, pr=pr, nh=nh, nt=nt, kg_pos=npy(kpos), kg_neg=npy(kneg),
               kg_loss=npy(kloss), margin=1.0,
               **{k.replace("grad_", "kggrad_"): v for k, v in grads_of(m).items()})
    qu = rng.randint(0, U, 3)
    torch.manual_seed(s3)
    out.update(qu=qu, eval_rec=npy(m.evaluateRec(LT(qu))))
    q, qr = rng.randint(0, E, 4), rng.randint(0, R, 4)
    out.update(q=q, qr=qr, eval_head=npy(m.evaluateHead(LT(q), LT(qr))),
               eval_tail=npy(m.evaluateTail(LT(q), LT(qr))))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def ranking_case(name, seed):
    """getRecPerformance / getKGPerformance (utils/misc.py) on tie-free rows."""
    rng = np.random.RandomState(seed)
    n, topn = 60, 10
    out = {"topn": topn}
    for c in range(6):
        scores = rng.permutation(n).astype(np.float32) / 7.0     # distinct values: no ties
        gold = set(int(x) for x in rng.choice(n, 4, replace=False))
        filt = set(int(x) for x in rng.choice(n, 9, replace=False)) - gold if c % 2 == 0 else None
        f1, p, r, hit, ndcg, top_ids = getRecPerformance(scores, gold, fliter_samples=filt, topn=topn)
        hits, ranks, gold_ids = getKGPerformance(scores, gold, fliter_samples=filt, topn=topn)
        out.update({f"c{c}_scores": scores, f"c{c}_gold": np.array(sorted(gold)),
                    f"c{c}_filter": np.array(sorted(filt)) if filt is not None else np.array([-1]),
                    f"c{c}_rec": np.array([f1, p, r, hit, ndcg]), f"c{c}_top_ids": np.array(top_ids),
                    f"c{c}_kg_hits": np.array(hits), f"c{c}_kg_ranks": np.array(ranks),
                    f"c{c}_kg_gold_ids": np.array(gold_ids)})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def regs_case(name, seed):
    """The drivers' regularisers (utils/loss.py:18-23) on rows on both sides of the unit sphere:
    values and autograd gradients."""
    torch.manual_seed(seed)
    ent = torch.randn(B * 3, D) * torch.linspace(0.12, 0.34, B * 3).view(-1, 1)      # |row|^2 from ~0.3 to ~2.3
    rel = torch.randn(B, D) * 0.25
    nrm = torch.randn(B, D) * 0.22
    ent.requires_grad_(True); rel.requires_grad_(True); nrm.requires_grad_(True)
    nl_e = ref_loss.normLoss(ent)
    nl_r = ref_loss.normLoss(rel)
    ol = ref_loss.orthogonalLoss(rel, nrm)
    (nl_e + nl_r + ol).backward()
    out = {"ent": npy(ent), "rel": npy(rel), "norm": npy(nrm), "norm_loss_ent": npy(nl_e), "norm_loss_rel": npy(nl_r),
           "orth_loss": npy(ol), "grad_ent": npy(ent.grad), "grad_rel": npy(rel.grad), "grad_norm": npy(nrm.grad)}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "regs":          # add one file without rewriting the others
        regs_case("regularisers", 12)
        return
    seed = 7
    for l1 in (False, True):
        tag = "l1" if l1 else "l2"
        kg_case(f"transe_{tag}", TransEModel, l1, seed)
        kg_case(f"transh_{tag}", TransHModel, l1, seed + 1)
        kg_case(f"transr_{tag}", TransRModel, l1, seed + 2)
        for gumbel in (False, True):
            gt = "gumbel" if gumbel else "soft"
            tup_case(f"transup_{tag}_{gt}", l1, gumbel, seed + 3)
            ktup_case(f"jtransup_{tag}_{gt}", l1, gumbel, seed + 4)
    ranking_case("ranking", seed + 5)
    regs_case("regularisers", 12)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
