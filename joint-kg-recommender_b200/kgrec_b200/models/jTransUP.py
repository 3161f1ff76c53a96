"""KTUP (`-model_type jtransup`) on the CUDA engine.  Mirrors jTransUP/models/jTransUP.py
(constructor 25-112, paddingItems 114-120, forward 122-161, evaluateRec 163-191,
evaluateHead/Tail 193-247, getPreferences 250-260)."""
import torch
import torch.nn as nn

from .. import _lib
from .. import functional as KF
from .base import _make_tables
from .transUP import RecModelBase


def build_model(FLAGS, user_total, item_total, entity_total, relation_total, i_map=None, e_map=None, new_map=None):
    return jTransUPModel(L1_flag=FLAGS.L1_flag, embedding_size=FLAGS.embedding_size, user_total=user_total,
                         item_total=item_total, entity_total=entity_total, relation_total=relation_total,
                         i_map=i_map, new_map=new_map, isShare=FLAGS.share_embeddings,
                         use_st_gumbel=FLAGS.use_st_gumbel)


def build_item2ent(item_total, pad, i_map, new_map):
    """paddingItems (jTransUP.py:114-120) for every item at once: item -> new_map[i_map[item]][0], unaligned
    (-1) or unmapped items -> the padding row.  Vectorised: the dicts are walked once by numpy.fromiter, not
    once per item per forward call."""
    import numpy as np
    table = np.full(item_total, pad, dtype=np.int32)
    if i_map is None or new_map is None or item_total == 0:
        return torch.from_numpy(table)
    n_i, n_m = len(i_map), len(new_map)
    items = np.fromiter(i_map.keys(), dtype=np.int64, count=n_i)
    idx = np.fromiter(i_map.values(), dtype=np.int64, count=n_i)
    mkeys = np.fromiter(new_map.keys(), dtype=np.int64, count=n_m)
    ments = np.fromiter((v[0] for v in new_map.values()), dtype=np.int64, count=n_m)
    ent_of = np.full(int(max(mkeys.max(initial=-1), idx.max(initial=-1))) + 1, -1, dtype=np.int64)
    ent_of[mkeys] = ments
    ok = (items >= 0) & (items < item_total)
    e = ent_of[idx[ok]]
    table[items[ok]] = np.where(e >= 0, e, pad).astype(np.int32)
    return torch.from_numpy(table)


class jTransUPModel(RecModelBase):
    MODEL = _lib.KTUP
    TABLES = {"user": "user_embeddings", "item": "item_embeddings", "ent": "ent_embeddings",
              "rel": "rel_embeddings", "norm": "norm_embeddings",
              "pref": "pref_embeddings", "pref_norm": "pref_norm_embeddings"}

    def __init__(self, L1_flag, embedding_size, user_total, item_total, entity_total, relation_total,
                 i_map, new_map, isShare, use_st_gumbel):
        super().__init__()
        self.L1_flag = L1_flag
        self.is_share = isShare
        self.use_st_gumbel = use_st_gumbel
        self.embedding_size = embedding_size
        self.user_total = user_total
        self.item_total = item_total
        self.ent_total = entity_total + 1          # + zero padding row for unaligned items
        self.rel_total = relation_total
        self.i_map = i_map
        self.new_map = new_map
        d = embedding_size
        # two groups, as the reference draws them: the TUP tables (jTransUP.py:52-66), then TransH's (83-94)
        _make_tables(self, [("user_embeddings", user_total, d, True), ("item_embeddings", item_total, d, True),
                            ("pref_embeddings", relation_total, d, True),
                            ("pref_norm_embeddings", relation_total, d, True)])
        _make_tables(self, [("ent_embeddings", entity_total, d, True, 1, {"padding_idx": self.ent_total - 1}),
                            ("rel_embeddings", relation_total, d, True), ("norm_embeddings", relation_total, d, True)])
        # paddingItems (jTransUP.py:114-120) as a device lookup table built once:
        # item -> aligned entity row, unaligned -> the padding row
        self.register_buffer("item2ent", build_item2ent(item_total, self.ent_total - 1, i_map, new_map), persistent=False)
        self._finish_init()

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._item2ent = self.item2ent           # follow .cuda() / .to()
        return out

    def paddingItems(self, i_ids, pad_index):
        """Reference helper kept for API parity; the kernels use the device table."""
        t = self.item2ent.cpu()
        return [int(t[int(i)]) for i in i_ids]

    def _mix_tables(self):
        return (self.pref_embeddings.weight + self.rel_embeddings.weight,
                self.pref_norm_embeddings.weight + self.norm_embeddings.weight, 0.5)

    def _pair_vectors(self, u_id, i_ids):
        u, i_e = super()._pair_vectors(u_id, i_ids)
        dev = self.device
        e = self.ent_embeddings.weight[self.item2ent[KF.as_index(i_ids, dev).long()].long()]
        return u, i_e + e

    def forward(self, ratings, triples, is_rec=True, gumbel_u=None):
        if is_rec and ratings is not None:
            u_ids, i_ids = ratings
            return self._score(_lib.KTUP, u_ids, i_ids, None, gumbel_u)
        if not is_rec and triples is not None:
            h, t, r = triples
            return self._score(_lib.TRANSH, h, t, r)         # jTransUP.py:144-157
        raise NotImplementedError

    def kg_rank_loss(self, pos, neg, margin=1.0, loss="margin", batch_pos=None):
        """Fused KG branch: TransH on (ent, rel, norm)."""
        return self._rank_loss(_lib.TRANSH, pos, neg, loss, margin, batch_pos)

    def kg_rank_loss_corrupt(self, pos, corrupt, margin=1.0, loss="margin", batch_pos=None):
        return self._rank_loss_corrupt(_lib.TRANSH, pos, corrupt, loss, margin, batch_pos)

    def kg_loss_step_corrupt(self, pos, corrupt, margin=1.0, loss="margin", batch_pos=None, grad_loss=1.0, reg=False):
        """KG branch forward + loss + backward in one kernel (the TransH group kernel on ent / rel / norm);
        grad_loss carries kg_lambda, reg the regularisers (knowledgable_recommendation.py:368-383)."""
        return self._loss_step_corrupt(_lib.TRANSH, pos, corrupt, loss, margin, batch_pos, grad_loss, reg)

    # -- evaluation ----------------------------------------------------------------------------
    def _rec_catalog(self):
        """ie = Item + Ent[item2ent] for every item (jTransUP.py:177-181), built on the device."""
        import ctypes as C
        dev = self._require_cuda()
        T = KF.make_tables(self._weights(), self.embedding_size, self.L1_flag, self.use_st_gumbel, self._item2ent)
        out = torch.empty((self.item_total, self.embedding_size), dtype=torch.float32, device=dev)
        lib = _lib.load()
        _lib.check(lib.kgrec_ktup_item_table(C.byref(T), 0, self.item_total, C.c_void_p(out.data_ptr()),
                                             out.stride(0), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        KF.count_launches(1)
        return out

    def _sub_catalog_unsupported(self, ids):
        # The reference honours all_i_ids / all_e_ids only when is_share (jTransUP.py:166-181, 196-199); -model_type
        # jtransup forces share_embeddings False (base.py:122-123), so the stock drivers never get here.
        if self.is_share and ids is not None:
            raise NotImplementedError("jTransUPModel(isShare=True): evaluation on a sub-catalog (all_i_ids / all_e_ids) "
                                      "is not built; the reference's drivers always run jtransup with isShare=False")

    def evaluateRec(self, u_ids, all_i_ids=None, gumbel_u=None):
        self._sub_catalog_unsupported(all_i_ids)
        return self._rec_scores(u_ids, gumbel_u)

    def _ent_catalog(self):
        return self.ent_embeddings.weight.detach()      # includes the padding row (jTransUP.py:195)

    def evaluateHead(self, t, r, all_e_ids=None):
        self._sub_catalog_unsupported(all_e_ids)
        return self._eval(_lib.TRANSH, _lib.SIDE_HEAD, t, r, "scores", catalog=self._ent_catalog())

    def evaluateTail(self, h, r, all_e_ids=None):
        self._sub_catalog_unsupported(all_e_ids)
        return self._eval(_lib.TRANSH, _lib.SIDE_TAIL, h, r, "scores", catalog=self._ent_catalog())

    def topk(self, side, q, r, k=10, filter_csr=None, catalog=None, id_base=0):
        s = _lib.SIDE_HEAD if side == "head" else _lib.SIDE_TAIL
        cat = self._ent_catalog() if catalog is None else catalog
        return self._eval(_lib.TRANSH, s, q, r, "topk", catalog=cat, id_base=id_base, k=k, filter_csr=filter_csr)
