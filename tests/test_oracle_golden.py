"""Pin the numpy oracle (oracle/kg_oracle.py) against golden vectors recorded from
the unmodified reference classes (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import kg_oracle as O

RTOL, ATOL = 2e-5, 2e-6   # fp32 re-association noise between torch and numpy


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


KG = {
    "transe": (lambda g: (g["w_ent_embeddings"], g["w_rel_embeddings"]), O.transe_score, O.transe_eval, O.transe_grads),
    "transh": (lambda g: (g["w_ent_embeddings"], g["w_rel_embeddings"], g["w_norm_embeddings"]),
               O.transh_score, O.transh_eval, O.transh_grads),
    "transr": (lambda g: (g["w_ent_embeddings"], g["w_rel_embeddings"], g["w_proj_embeddings"]),
               O.transr_score, O.transr_eval, O.transr_grads),
}
GRAD_KEYS = {"ent": "ent_embeddings", "rel": "rel_embeddings", "norm": "norm_embeddings",
             "proj": "proj_embeddings", "user": "user_embeddings", "item": "item_embeddings",
             "pref": "pref_embeddings", "pref_norm": "pref_norm_embeddings"}


@pytest.mark.parametrize("model", ["transe", "transh", "transr"])
@pytest.mark.parametrize("tag", ["l2", "l1"])
def test_kg_models(golden, model, tag):
    g = golden(f"{model}_{tag}")
    tables, score, evalf, grads = KG[model]
    T = tables(g)
    l1 = bool(g["l1"])
    pos = score(*T, g["ph"], g["pt"], g["pr"], l1)
    neg = score(*T, g["nh"], g["nt"], g["nr"], l1)
    close(pos, g["pos"])
    close(neg, g["neg"])
    close(O.margin_loss(pos, neg, float(g["margin"])), g["loss"])
    close(evalf(*T, g["q"], g["qr"], l1, "head"), g["eval_head"])
    close(evalf(*T, g["q"], g["qr"], l1, "tail"), g["eval_tail"])
    # d(sum scores)/d tables
    gs = grads(*T, g["ph"], g["pt"], g["pr"], l1, np.ones(len(g["ph"]), np.float32))
    for k, v in gs.items():
        close(v, g["gsum_" + GRAD_KEYS[k]], rtol=1e-4, atol=1e-5)
    # d(margin loss)/d tables = pos grads with dL/dpos + neg grads with dL/dneg
    gp, gn = O.margin_loss_grads(pos, neg, float(g["margin"]))
    a = grads(*T, g["ph"], g["pt"], g["pr"], l1, gp)
    b = grads(*T, g["nh"], g["nt"], g["nr"], l1, gn)
    for k in a:
        close(a[k] + b[k], g["grad_" + GRAD_KEYS[k]], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag", ["l2", "l1"])
@pytest.mark.parametrize("mode", ["soft", "gumbel"])
def test_transup(golden, tag, mode):
    g = golden(f"transup_{tag}_{mode}")
    T = (g["w_user_embeddings"], g["w_item_embeddings"], g["w_pref_embeddings"], g["w_pref_norm_embeddings"])
    l1 = bool(g["l1"])
    npos = g.get("noise_pos")
    nneg = g.get("noise_neg")
    nev = g.get("noise_eval")
    pos = O.tup_score(*T, g["u"], g["pi"], l1, npos)
    neg = O.tup_score(*T, g["u"], g["ni"], l1, nneg)
    close(pos, g["pos"])
    close(neg, g["neg"])
    close(O.bpr_loss(pos, neg, float(g["target"])), g["loss"])
    close(O.tup_eval(*T, g["qu"], l1, nev), g["eval"], rtol=1e-4, atol=1e-5)
    gs = O.tup_grads(*T, g["u"], g["pi"], l1, np.ones(len(g["u"]), np.float32), npos)
    for k, v in gs.items():
        close(v, g["gsum_" + GRAD_KEYS[k]], rtol=2e-4, atol=2e-5)
    gp, gn = O.bpr_loss_grads(pos, neg, float(g["target"]))
    a = O.tup_grads(*T, g["u"], g["pi"], l1, gp, npos)
    b = O.tup_grads(*T, g["u"], g["ni"], l1, gn, nneg)
    for k in a:
        close(a[k] + b[k], g["grad_" + GRAD_KEYS[k]], rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("tag", ["l2", "l1"])
@pytest.mark.parametrize("mode", ["soft", "gumbel"])
def test_jtransup(golden, tag, mode):
    g = golden(f"jtransup_{tag}_{mode}")
    T = (g["w_user_embeddings"], g["w_item_embeddings"], g["w_ent_embeddings"], g["w_rel_embeddings"],
         g["w_norm_embeddings"], g["w_pref_embeddings"], g["w_pref_norm_embeddings"], g["item2ent"])
    l1 = bool(g["l1"])
    npos, nneg, nev = g.get("noise_pos"), g.get("noise_neg"), g.get("noise_eval")
    assert T[2].shape[0] == 38 and not T[2][-1].any()          # E+1 rows, zero padding row
    pos = O.ktup_rec_score(*T, g["u"], g["pi"], l1, npos)
    neg = O.ktup_rec_score(*T, g["u"], g["ni"], l1, nneg)
    close(pos, g["pos"])
    close(neg, g["neg"])
    close(O.bpr_loss(pos, neg, float(g["target"])), g["loss"])
    close(O.ktup_rec_eval(*T, g["qu"], l1, nev), g["eval_rec"], rtol=1e-4, atol=1e-5)
    gp, gn = O.bpr_loss_grads(pos, neg, float(g["target"]))
    a = O.ktup_rec_grads(*T, g["u"], g["pi"], l1, gp, npos)
    b = O.ktup_rec_grads(*T, g["u"], g["ni"], l1, gn, nneg)
    for k in a:
        close(a[k] + b[k], g["grad_" + GRAD_KEYS[k]], rtol=2e-4, atol=2e-6)
    # KG branch == TransH on the KTUP tables (jTransUP.py:144-157, 193-247)
    H = (T[2], T[3], T[4])
    kpos = O.transh_score(*H, g["ph"], g["pt"], g["pr"], l1)
    kneg = O.transh_score(*H, g["nh"], g["nt"], g["pr"], l1)
    close(kpos, g["kg_pos"])
    close(kneg, g["kg_neg"])
    close(O.margin_loss(kpos, kneg, float(g["margin"])), g["kg_loss"])
    close(O.transh_eval(*H, g["q"], g["qr"], l1, "head"), g["eval_head"])
    close(O.transh_eval(*H, g["q"], g["qr"], l1, "tail"), g["eval_tail"])
    gp, gn = O.margin_loss_grads(kpos, kneg, float(g["margin"]))
    a = O.transh_grads(*H, g["ph"], g["pt"], g["pr"], l1, gp)
    b = O.transh_grads(*H, g["nh"], g["nt"], g["pr"], l1, gn)
    for k in a:
        close(a[k] + b[k], g["kggrad_" + GRAD_KEYS[k]], rtol=1e-4, atol=1e-5)
    for k in ("user", "item", "pref", "pref_norm"):          # untouched by the KG branch
        assert not g["kggrad_" + GRAD_KEYS[k]].any()


def test_ranking(golden):
    g = golden("ranking")
    topn = int(g["topn"])
    for c in range(6):
        scores = g[f"c{c}_scores"]
        gold = set(int(x) for x in g[f"c{c}_gold"])
        f = g[f"c{c}_filter"]
        filt = None if (len(f) == 1 and f[0] == -1) else set(int(x) for x in f)
        top = O.rec_topk(scores, filt, topn)
        assert top == [int(x) for x in g[f"c{c}_top_ids"]]
        close(np.array(O.rec_metrics(top, gold)), g[f"c{c}_rec"], rtol=1e-12, atol=0)
        ranks = O.kg_ranks(scores, gold, filt, topn)
        ids = [int(x) for x in g[f"c{c}_kg_gold_ids"]]
        assert sorted(ranks) == sorted(ids)
        assert [ranks[i][1] for i in ids] == [int(x) for x in g[f"c{c}_kg_ranks"]]
        assert [ranks[i][0] for i in ids] == [int(x) for x in g[f"c{c}_kg_hits"]]


def test_ndcg_reference_doctests():
    """Known-answer values in the reference's own doctests: utils/evaluation.py:47-59, 86-97."""
    r = [3, 2, 3, 0, 0, 1, 2, 2, 3, 0]
    assert O.dcg_at_k(r, 1) == 3.0 and O.dcg_at_k(r, 1, method=1) == 3.0
    assert O.dcg_at_k(r, 2) == 5.0
    assert O.dcg_at_k(r, 2, method=1) == pytest.approx(4.2618595071429155, rel=1e-15)
    assert O.dcg_at_k(r, 10) == pytest.approx(9.6051177391888114, rel=1e-15)
    assert O.dcg_at_k(r, 11) == pytest.approx(9.6051177391888114, rel=1e-15)
    assert O.ndcg_at_k(r, 1) == 1.0
    assert O.ndcg_at_k([2, 1, 2, 0], 4) == pytest.approx(0.9203032077642922, rel=1e-15)
    assert O.ndcg_at_k([2, 1, 2, 0], 4, method=1) == pytest.approx(0.96519546960144276, rel=1e-15)
    assert O.ndcg_at_k([0], 1) == 0.0 and O.ndcg_at_k([1], 2) == 1.0


def test_invariants():
    """SURVEY section 4 invariants on the oracle itself (fp64)."""
    rng = np.random.RandomState(0)
    d, E, R, B = 12, 30, 4, 6
    ent, rel, nrm = rng.randn(E, d), rng.randn(R, d), rng.randn(R, d)
    h, t, r = rng.randint(0, E, B), rng.randint(0, E, B), rng.randint(0, R, B)
    for l1 in (False, True):
        s = O.transe_score(ent, rel, h, t, r, l1)
        close(O.transe_eval(ent, rel, h, r, l1, "tail")[np.arange(B), t], s, 1e-12, 1e-12)
        close(O.transe_eval(ent, rel, t, r, l1, "head")[np.arange(B), h], s, 1e-12, 1e-12)
        sh = O.transh_score(ent, rel, nrm, h, t, r, l1)
        close(O.transh_eval(ent, rel, nrm, h, r, l1, "tail")[np.arange(B), t], sh, 1e-10, 1e-10)
        close(O.transh_score(ent, rel, 0 * nrm, h, t, r, l1), s, 1e-12, 1e-12)
        eye = np.tile(np.eye(d).reshape(1, -1), (R, 1))
        close(O.transr_score(ent, rel, eye, h, t, r, l1), s, 1e-12, 1e-12)


@pytest.mark.parametrize("name,with_norm", [("transe", False), ("transh", True)])
@pytest.mark.parametrize("tag", ["l2", "l1"])
def test_torch_port_matches_golden(golden, name, with_norm, tag):
    """oracle/torch_port.py (the CPU-baseline arm of bench.py) against the same vectors."""
    import torch
    from oracle import torch_port as TP
    g = golden(f"{name}_{tag}")
    E, D = g["w_ent_embeddings"].shape
    m = TP.TransPort(bool(g["l1"]), D, E, g["w_rel_embeddings"].shape[0], with_norm)
    m.load_state_dict({k[2:] + ".weight": torch.from_numpy(v) for k, v in g.items() if k.startswith("w_")})
    LT = lambda x: torch.from_numpy(x).long()  # noqa: E731
    loss = TP.train_step(m, (LT(g["ph"]), LT(g["pt"]), LT(g["pr"])), (LT(g["nh"]), LT(g["nt"]), LT(g["nr"])))
    close(loss.item(), g["loss"])
    for k in ("ent", "rel") + (("norm",) if with_norm else ()):
        close(getattr(m, k + "_embeddings").weight.grad.numpy(), g[f"grad_{k}_embeddings"], rtol=1e-4, atol=1e-6)
    close(m.evaluate_side(LT(g["q"]), LT(g["qr"]), True).detach().numpy(), g["eval_head"])
    close(m.evaluate_side(LT(g["q"]), LT(g["qr"]), False).detach().numpy(), g["eval_tail"])


def test_regularisers_against_reference(golden):
    """normLoss / orthogonalLoss (utils/loss.py:18-23): values and autograd gradients recorded from the
    reference functions; the step kernels' fused regularisers (reg_flags) are checked against the same
    formulas on the GPU."""
    g = golden("regularisers")
    ent, rel, nrm = g["ent"], g["rel"], g["norm"]
    np.testing.assert_allclose(O.norm_loss(ent), g["norm_loss_ent"], rtol=1e-5)
    np.testing.assert_allclose(O.norm_loss(rel), g["norm_loss_rel"], rtol=1e-5)
    np.testing.assert_allclose(O.orthogonal_loss(rel, nrm), g["orth_loss"], rtol=1e-5)
    np.testing.assert_allclose(O.norm_loss_grads(ent), g["grad_ent"], rtol=1e-5, atol=1e-7)
    gr, gw = O.orthogonal_loss_grads(rel, nrm)
    np.testing.assert_allclose(gr + O.norm_loss_grads(rel), g["grad_rel"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gw, g["grad_norm"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", ["transe_l2", "transe_l1", "transh_l2"])
def test_oracle_at_config_shape_cfg2(golden, name):
    """The oracle at BASELINE configs[1] (d=100, |E|=100k, |R|=500, 1024 positives x 10 negatives) against vectors
    recorded from the reference classes (tests/golden/make_golden_cfg2.py); the tables are rebuilt from the seed by
    the drop-in constructors, which consume the generator as the reference's do (checked on the recorded rows)."""
    import torch
    import kgrec_b200 as K
    g = golden("cfg2_" + name)
    l1, transh = name.endswith("l1"), name.startswith("transh")
    torch.manual_seed(int(g["seed"]))
    m = (K.TransHModel if transh else K.TransEModel)(l1, 100, 100_000, 500).cpu()
    W = {k.replace("_embeddings.weight", ""): v.detach().numpy() for k, v in m.state_dict().items()}
    np.testing.assert_array_equal(W["ent"][:4], g["ent_rows_check"])
    np.testing.assert_array_equal(W["rel"][-4:], g["rel_rows_check"])
    T = (W["ent"], W["rel"]) + ((W["norm"],) if transh else ())
    score, grads = (O.transh_score, O.transh_grads) if transh else (O.transe_score, O.transe_grads)
    ph, pt, pr, c = g["ph"], g["pt"], g["pr"], g["corrupt"].astype(np.int64)
    K_ = c.size // ph.size
    head = c < 0
    cid = np.where(head, ~c, c)
    nh, nt, nr = np.where(head, cid, np.repeat(ph, K_)), np.where(head, np.repeat(pt, K_), cid), np.repeat(pr, K_)
    pos, neg = score(*T, ph, pt, pr, l1), score(*T, nh, nt, nr, l1)
    close(pos, g["pos_scores"], rtol=5e-5)
    close(neg, g["neg_scores"], rtol=5e-5)
    posr = np.repeat(pos, K_)
    close(O.margin_loss(posr, neg, 1.0), g["loss"], rtol=5e-5)
    gp, gn = O.margin_loss_grads(posr, neg, 1.0)
    a = grads(*T, np.repeat(ph, K_), np.repeat(pt, K_), nr, l1, gp)
    b = grads(*T, nh, nt, nr, l1, gn)
    ent = a["ent"] + b["ent"]
    if l1:      # sign(e) at residual components within rounding of zero may differ between torch and numpy
        bad = np.abs(ent[g["ent_grad_ids"]] - g["ent_grad_rows"]) > 1e-4
        assert bad.sum() <= 8
    else:
        close(ent[g["ent_grad_ids"]], g["ent_grad_rows"], rtol=2e-4, atol=2e-5)
        close(a["rel"] + b["rel"], g["rel_grad"], rtol=2e-4, atol=5e-5)
    close((ent.astype(np.float64) ** 2).sum(), g["ent_grad_sqnorm"], rtol=1e-3)
