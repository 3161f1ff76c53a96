"""TransR on the CUDA engine.  Mirrors jTransUP/models/transR.py (constructor 17-63,
forward 65-78, evaluateHead/Tail 80-128; projections utils/misc.py:21-33)."""
import torch

from .. import _lib
from .. import functional as KF
from .transE import KGModelBase


def build_model(FLAGS, user_total, item_total, entity_total, relation_total, i_map=None, e_map=None, new_map=None):
    return TransRModel(L1_flag=FLAGS.L1_flag, embedding_size=FLAGS.embedding_size,
                       ent_total=entity_total, rel_total=relation_total)


class TransRModel(KGModelBase):
    MODEL = _lib.TRANSR
    TABLES = {"ent": "ent_embeddings", "rel": "rel_embeddings", "proj": "proj_embeddings"}

    def __init__(self, L1_flag, embedding_size, ent_total, rel_total):
        super().__init__(L1_flag, embedding_size, ent_total, rel_total)
        self.max_entity_batch = 10
        self._finish_init()

    def _table_specs(self):
        # d x d matrix per relation stored as a row of d*d; xavier, NOT normalised (transR.py:42,55)
        d = self.embedding_size
        return super()._table_specs() + [("proj_embeddings", self.rel_total, d * d, False)]

    # Full-catalog TransR (the reference's one dense contraction, misc.py:29-33): queries are
    # grouped by relation, the catalog is projected once per distinct relation by a library
    # GEMM, and the distance part runs in the evaluation kernel on explicit query vectors.
    def _eval_side(self, side, q, r, mode, **kw):
        dev = self._require_cuda()
        q, r = KF.as_index(q, dev).long(), KF.as_index(r, dev).long()
        d = self.embedding_size
        ent = self.ent_embeddings.weight.detach()
        rel = self.rel_embeddings.weight.detach()
        proj = self.proj_embeddings.weight.detach()
        results = {}
        for rid in torch.unique(r).tolist():
            sel = (r == rid).nonzero().view(-1)
            m = proj[rid].view(d, d)
            cat = (ent @ m.t()).contiguous()                       # M e for every entity
            pq = ent[q[sel]] @ m.t()
            c = pq - rel[rid] if side == _lib.SIDE_HEAD else pq + rel[rid]
            qvec = torch.cat([c, torch.zeros_like(c)], dim=1).contiguous()
            sub = {k: (v[sel] if torch.is_tensor(v) and v.shape[:1] == r.shape else v) for k, v in kw.items()}
            results[rid] = (sel, self._eval(self.MODEL, side, None, None, mode, catalog=cat, qvec=qvec, **sub))
        first = next(iter(results.values()))[1]
        out = torch.empty((q.numel(),) + tuple(first.shape[1:]), dtype=first.dtype, device=dev)
        for sel, res in results.values():
            out[sel] = res
        return out

    def evaluateHead(self, t, r, all_e_ids=None):
        return self._eval_side(_lib.SIDE_HEAD, t, r, "scores")

    def evaluateTail(self, h, r, all_e_ids=None):
        return self._eval_side(_lib.SIDE_TAIL, h, r, "scores")

    def topk(self, side, q, r, k=10, filter_csr=None, catalog=None, id_base=0):
        if filter_csr is not None or catalog is not None:
            raise NotImplementedError("TransR top-K supports the whole, unfiltered entity table only")
        s = _lib.SIDE_HEAD if side == "head" else _lib.SIDE_TAIL
        return self._eval_side(s, q, r, "topk", k=k)

    def rank_counts(self, *a, **kw):
        raise NotImplementedError("TransR rank counts: use evaluateHead/evaluateTail")
