"""SURVEY 8f row 3 for TransR: `kgrec_b200.metrics.evaluate_kg` (on-chip rank counts + the filtered / multi-gold
correction on gathered rows) == the reference's argsort walk (knowledge_representation.py:28-105 ->
utils/misc.py:125-146, restated in oracle/kg_oracle.py) over the full score matrices of evaluateHead / evaluateTail
(transR.py:80-128), with filter sets, multi-gold queries, gold ids that are themselves filtered, and unsorted
relations (the native TransR evaluation groups the queries by relation and projects the catalog per run)."""
import numpy as np
import pytest
import torch

from oracle import kg_oracle as O

pytestmark = pytest.mark.gpu


def lt(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.long, device="cuda")


@pytest.mark.parametrize("d,l1", [(100, False), (64, True)])
def test_transr_driver_level_eval_matches_ranking_walk(d, l1):
    import kgrec_b200 as K
    from kgrec_b200 import metrics as KM
    torch.manual_seed(21 + d)
    rng = np.random.RandomState(21 + d)
    E, R, topn = 1800, 7, 10
    m = K.TransRModel(l1, d, E, R)

    def rand_dict(n_keys, lo, hi):
        out = {}
        while len(out) < n_keys:
            key = (int(rng.randint(0, E)), int(rng.randint(0, R)))
            out[key] = set(int(x) for x in rng.choice(E, rng.randint(lo, hi), replace=False))
        return out
    head_eval, tail_eval = rand_dict(70, 1, 4), rand_dict(85, 1, 4)
    head_all = [{k: set(int(x) for x in rng.choice(E, 15, replace=False)) for k in list(head_eval)[::2]}]
    tail_all = [{k: set(int(x) for x in rng.choice(E, 25, replace=False)) for k in list(tail_eval)[::3]}, {}]
    # one query whose filter set holds one of its own gold ids: the reference's walk never reaches it
    k0 = next(iter(tail_eval))
    tail_all[1][k0] = {next(iter(tail_eval[k0]))}
    got = KM.evaluate_kg(m, head_eval, tail_eval, head_all, tail_all, topn=topn, batch=32)
    want = {}
    for side, ev, alld in (("head", head_eval, head_all), ("tail", tail_eval, tail_all)):
        keys = list(ev)
        q, r = lt([k[0] for k in keys]), lt([k[1] for k in keys])
        full = (m.evaluateHead(q, r) if side == "head" else m.evaluateTail(q, r)).cpu().numpy()
        assert full.shape == (len(keys), E)
        res = []
        for b, key in enumerate(keys):
            filt = set()
            for dct in alld:
                if key in dct:
                    filt.update(dct[key])
            res.extend(O.kg_ranks(full[b], ev[key], filt, topn).values())
        want[side] = np.asarray(res, dtype=np.float64)
    for g_, side in ((got[2], "head"), (got[3], "tail")):
        np.testing.assert_allclose(g_, want[side].mean(axis=0), rtol=1e-12)
    tot = len(want["head"]) + len(want["tail"])
    np.testing.assert_allclose(got[0], (want["head"][:, 0].sum() + want["tail"][:, 0].sum()) / tot, rtol=1e-12)
    np.testing.assert_allclose(got[1], (want["head"][:, 1].sum() + want["tail"][:, 1].sum()) / tot, rtol=1e-12)
    m.check_indices()
