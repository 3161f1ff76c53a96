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
    EVAL_MAX_DIM = 128          # k_transr_project keeps M_r^T (d x d) and a 128-row tile in shared memory
    TABLES = {"ent": "ent_embeddings", "rel": "rel_embeddings", "proj": "proj_embeddings"}

    def __init__(self, L1_flag, embedding_size, ent_total, rel_total):
        super().__init__(L1_flag, embedding_size, ent_total, rel_total)
        self.max_entity_batch = 10
        self._finish_init()

    def _table_specs(self):
        # d x d matrix per relation stored as a row of d*d; xavier, NOT normalised (transR.py:42,55)
        d = self.embedding_size
        return super()._table_specs() + [("proj_embeddings", self.rel_total, d * d, False)]

    # Full-catalog TransR (transR.py:80-128; misc.py:29-33).  The reference projects the whole entity table with
    # every query's matrix; here the queries are sorted by relation (one device sort; the handful of run
    # boundaries is the only thing read back), and per distinct relation the library projects the catalog
    # once with a hand-written FP32 kernel and runs the distance / top-K / rank-count kernels on the projected
    # rows (csrc/eval_transr.cu).  Results come back in the caller's query order.
    def _runs(self, q, r):
        dev = self._require_cuda()
        q, r = KF.as_index(q, dev).long(), KF.as_index(r, dev).long()
        r_sorted, order = torch.sort(r, stable=True)
        rels, counts = torch.unique_consecutive(r_sorted, return_counts=True)
        host = torch.stack([rels, counts]).cpu()                 # the one device -> host read: a few run boundaries
        begin = torch.zeros(host.shape[1] + 1, dtype=torch.int64)
        begin[1:] = torch.cumsum(host[1], 0)
        return q[order].contiguous(), r_sorted.contiguous(), order, begin.contiguous(), host[0].contiguous()

    def _transr_call(self, fn_name, side, q, r, catalog, id_base, *mode_args, cat_ids=None):
        import ctypes as C
        dev = self._require_cuda()
        lib = _lib.load()
        cat = self._catalog() if catalog is None else catalog
        qs, rs, order, begin, rels = self._runs(q, r)
        nq, d = qs.numel(), self.embedding_size
        T = KF.make_tables(self._weights(), d, self.L1_flag)
        ws = torch.empty(int(lib.kgrec_transr_workspace_floats(nq, cat.shape[0], d)), dtype=torch.float32, device=dev)
        common = [C.byref(T), side, KF._ptr(qs), KF._ptr(rs), 8, nq, C.c_void_p(begin.data_ptr()), C.c_void_p(rels.data_ptr()),
                  rels.numel(), KF._ptr(cat), cat.stride(0), cat.shape[0], id_base]
        self._keep = (begin, rels, qs, rs)      # host / device arrays the raw pointers in `common` refer to
        return lib, common, ws, order, nq, cat, dev

    def _scores(self, side, q, r, catalog=None, id_base=0, cat_ids=None):
        lib, common, ws, order, nq, cat, dev = self._transr_call("scores", side, q, r, catalog, id_base)
        if nq == 0:
            return torch.zeros((0, cat.shape[0]), dtype=torch.float32, device=dev)
        res = torch.empty((nq, cat.shape[0]), dtype=torch.float32, device=dev)
        ids = cat_ids.to(torch.int32).contiguous() if cat_ids is not None else None
        _lib.check(lib.kgrec_transr_eval_scores(*common, KF._ptr(ids), KF._ptr(ws), KF._ptr(res), res.stride(0),
                                                KF._ptr(self._status_buf(dev)), KF._stream()))
        out = torch.empty_like(res)
        out[order] = res
        self.check_indices()
        return out

    def evaluateHead(self, t, r, all_e_ids=None):
        return self._scores(_lib.SIDE_HEAD, t, r)

    def evaluateTail(self, h, r, all_e_ids=None):
        return self._scores(_lib.SIDE_TAIL, h, r)

    def topk(self, side, q, r, k=10, filter_csr=None, catalog=None, id_base=0):
        """K best entities per query (uint64 keys), filter_csr in the caller's query order."""
        s = _lib.SIDE_HEAD if side == "head" else _lib.SIDE_TAIL
        lib, common, ws, order, nq, cat, dev = self._transr_call("topk", s, q, r, catalog, id_base)
        if nq == 0:
            return torch.zeros((0, k), dtype=torch.int64, device=dev)
        fptr = fids = None
        if filter_csr is not None:                       # permute the CSR rows into the sorted query order
            ptr, ids = filter_csr
            lens = (ptr[1:] - ptr[:-1])[order]
            fptr = torch.zeros(nq + 1, dtype=torch.int64, device=dev)
            fptr[1:] = torch.cumsum(lens, 0)
            src = torch.repeat_interleave(ptr[:-1][order] - fptr[:-1], lens) + torch.arange(int(fptr[-1]), device=dev)
            fids = ids[src].contiguous() if src.numel() else ids
        keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
        tws = torch.empty(int(lib.kgrec_eval_workspace_bytes(nq, k)) // 8 + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.kgrec_transr_eval_topk(*common, KF._ptr(ws), k, KF._ptr(fptr), KF._ptr(fids), KF._ptr(keys), KF._ptr(tws),
                                              tws.numel() * 8, KF._ptr(self._status_buf(dev)), KF._stream()))
        out = torch.empty_like(keys)
        out[order] = keys
        return out

    def gold_scores(self, side, q, r, gold_ids):
        """Scores of (query, gold) pairs by the catalog pass's own arithmetic (bit-identical to the scores
        rank_counts compares them with): the gathered gold rows as a sub-catalog, diagonal of the result."""
        dev = self._require_cuda()
        g = KF.as_index(gold_ids, dev).long()
        s = _lib.SIDE_HEAD if side == "head" else _lib.SIDE_TAIL
        q, r = KF.as_index(q, dev), KF.as_index(r, dev)
        out = torch.empty(g.numel(), dtype=torch.float32, device=dev)
        for lo in range(0, g.numel(), 512):
            hi = min(g.numel(), lo + 512)
            rows = self._catalog()[g[lo:hi]].contiguous()
            out[lo:hi] = self._scores(s, q[lo:hi], r[lo:hi], catalog=rows, cat_ids=g[lo:hi]).diagonal()
        return out

    def rank_counts(self, side, q, r, gold_ids, gold_scores=None, catalog=None, id_base=0, out=None):
        """#entities ranked strictly before each query's gold id ((score, id) order)."""
        s = _lib.SIDE_HEAD if side == "head" else _lib.SIDE_TAIL
        if gold_scores is None:
            gold_scores = self.gold_scores(side, q, r, gold_ids)
        lib, common, ws, order, nq, cat, dev = self._transr_call("rank", s, q, r, catalog, id_base)
        counts = torch.zeros(nq, dtype=torch.int32, device=dev)
        if nq:
            gs = gold_scores.to(dev, torch.float32)[order].contiguous()
            gi = KF.as_index(gold_ids, dev).to(torch.int32)[order].contiguous()
            _lib.check(lib.kgrec_transr_eval_rank_count(*common, KF._ptr(ws), KF._ptr(gs), KF._ptr(gi), KF._ptr(counts),
                                                        KF._ptr(self._status_buf(dev)), KF._stream()))
        res = torch.empty_like(counts)
        res[order] = counts
        if out is not None:
            out += res
            return out
        return res
