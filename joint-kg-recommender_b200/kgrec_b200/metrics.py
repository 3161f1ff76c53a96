"""Driver-level evaluation on the reduced kernel outputs.

The reference's evaluate() functions copy a [B, N] score matrix to the host for every batch,
fork `num_processes` workers per batch and argsort each row (item_recommendation.py:27-75,
knowledge_representation.py:28-105, utils/misc.py:61-248).  These two functions produce the
same numbers -- mean F1 / precision / recall / hit / NDCG@n for recommendation, hit@n and mean
filtered rank for KG completion -- from the on-chip top-K lists and rank counts, so that only
K ids (or one count) per query ever leave the GPU.

Semantics restated from the reference (ties broken by (score, id), see oracle/kg_oracle.py):
  rec : top-n ids of the user after dropping the filter set (train + other eval files),
        misc.py:213-248; users with an empty gold set are skipped (misc.py:169).
  KG  : for every gold id g of a query, rank(g) = #{ e not in filter, not gold : e sorts before g };
        hit = rank < topn (misc.py:125-146).  A gold id that is itself in the filter set is never
        reached by the reference's walk and is skipped here too.
"""
import numpy as np
import torch

from . import evaluation as KE
from . import functional as KF


def evaluate_rec(model, eval_dict, all_dicts=None, topn=10, batch=4096):
    """Mean (f1, precision, recall, hit, ndcg) over the users of eval_dict.

    eval_dict: {user: set(gold items)}; all_dicts: dicts whose items are filtered per user
    (the drivers pass [train_dict] + the other eval files' dicts, item_recommendation.py:108-111).
    """
    dev = model._require_cuda()
    users = [u for u, gold in eval_dict.items() if len(gold) > 0]
    rows = []
    for lo in range(0, len(users), batch):
        chunk = users[lo:lo + batch]
        csr = KE.build_filter_csr(chunk, all_dicts, dev) if all_dicts else None
        keys = model.topk_items(torch.tensor(chunk, dtype=torch.int64, device=dev), k=topn, filter_csr=csr)
        ids, _ = KE.keys_to_ids_scores(keys)
        rows.extend(KE.rec_metrics_from_topk(ids.cpu().tolist(), [eval_dict[u] for u in chunk]))
    if not rows:
        return (0.0,) * 5
    return tuple(float(x) for x in np.asarray(rows, dtype=np.float64).mean(axis=0))


def _kg_side(model, side, eval_dict, all_dicts, topn, batch):
    """[(hit, rank)] for every (query, gold id) of one side.  eval_dict: {(q, r): set(gold)}."""
    dev = model._require_cuda()
    from . import _lib
    kg = _lib.TRANSH if model.MODEL == _lib.KTUP else model.MODEL
    sd = _lib.SIDE_HEAD if side == "head" else _lib.SIDE_TAIL
    transr = model.MODEL == _lib.TRANSR

    def sub_scores(qq, rr, rows, ids):
        """[len(qq), len(ids)] scores against gathered catalog rows, by the catalog pass's own arithmetic (bit-identical
        to the scores the rank-count kernel compares).  TransR projects the gathered rows per distinct relation
        (TransRModel._scores -> csrc/eval_transr.cu)."""
        if transr:
            return model._scores(sd, qq, rr, catalog=rows, cat_ids=ids)
        return model._eval(kg, sd, qq, rr, "scores", catalog=rows, cat_ids=ids)
    queries = [k for k, gold in eval_dict.items() if len(gold) > 0]
    results = []
    catalog = model.ent_embeddings.weight.detach()
    for lo in range(0, len(queries), batch):
        chunk = queries[lo:lo + batch]
        # one row per (query, gold id): the rank-count kernel takes one gold per query row
        qid, gold, excl = [], [], []
        for qi, key in enumerate(chunk):
            g_all = eval_dict[key]
            filt = set()
            for d in all_dicts or ():
                if key in d:
                    filt.update(d[key])
            for g in g_all:
                if g in filt:
                    continue                      # the reference's walk skips filtered ids before testing gold
                qid.append(qi)
                gold.append(g)
                excl.append(sorted((filt | g_all) - {g}))
        if not gold:
            continue
        q = torch.tensor([chunk[i][0] for i in qid], dtype=torch.int64, device=dev)
        r = torch.tensor([chunk[i][1] for i in qid], dtype=torch.int64, device=dev)
        gt = torch.tensor(gold, dtype=torch.int64, device=dev)
        # gold scores from the evaluation kernel itself (bit-identical to the catalog pass)
        gs = torch.empty(gt.numel(), dtype=torch.float32, device=dev)
        for glo in range(0, gt.numel(), 512):
            ghi = min(gt.numel(), glo + 512)
            gs[glo:ghi] = sub_scores(q[glo:ghi], r[glo:ghi], catalog[gt[glo:ghi]].contiguous(), gt[glo:ghi]).diagonal()
        if transr:
            counts = model.rank_counts(side, q, r, gt, gold_scores=gs).to(torch.int64)
        else:
            counts = model._eval(kg, sd, q, r, "rank", catalog=catalog, gold_scores=gs, gold_ids=gt).to(torch.int64)
        # correction: filtered ids and the other gold ids that sort before the gold do not count.
        # Their scores come from the same evaluation kernel on the gathered rows (bit-identical).
        flat_row = torch.tensor([i for i, e in enumerate(excl) for _ in e], dtype=torch.int64, device=dev)
        flat_ids = torch.tensor([x for e in excl for x in e], dtype=torch.int64, device=dev)
        if flat_ids.numel():
            uniq, inv = torch.unique(flat_ids, return_inverse=True)
            sub = catalog[uniq].contiguous()
            before = torch.zeros_like(counts)
            for qlo in range(0, q.numel(), 2048):          # [rows, n_unique] score blocks
                qhi = min(q.numel(), qlo + 2048)
                sel = (flat_row >= qlo) & (flat_row < qhi)
                if not bool(sel.any()):
                    continue
                m = sub_scores(q[qlo:qhi], r[qlo:qhi], sub, uniq)
                rr, cc = flat_row[sel] - qlo, inv[sel]
                s_e = m[rr, cc]
                s_g, i_g = gs[flat_row[sel]], gt[flat_row[sel]]
                lt = (s_e < s_g) | ((s_e == s_g) & (flat_ids[sel] < i_g))
                before.index_add_(0, flat_row[sel], lt.to(torch.int64))
            counts = counts - before
        ranks = counts.cpu().tolist()
        results.extend((1 if rk < topn else 0, rk) for rk in ranks)
    return results


def evaluate_kg(model, eval_head_dict, eval_tail_dict, all_head_dicts=None, all_tail_dicts=None, topn=10, batch=2048):
    """(avg_hit, avg_mean_rank, head (hit, rank), tail (hit, rank)) as knowledge_representation.py:66-87
    logs them.  eval_head_dict: {(t, r): set(gold heads)}, eval_tail_dict: {(h, r): set(gold tails)}."""
    head = _kg_side(model, "head", eval_head_dict, all_head_dicts, topn, batch)
    tail = _kg_side(model, "tail", eval_tail_dict, all_tail_dicts, topn, batch)
    h = np.asarray(head, dtype=np.float64).mean(axis=0) if head else np.zeros(2)
    t = np.asarray(tail, dtype=np.float64).mean(axis=0) if tail else np.zeros(2)
    n_h, n_t = len(head), len(tail)
    tot = max(1, n_h + n_t)
    return (float(h[0] * n_h + t[0] * n_t) / tot, float(h[1] * n_h + t[1] * n_t) / tot,
            (float(h[0]), float(h[1])), (float(t[0]), float(t[1])))
