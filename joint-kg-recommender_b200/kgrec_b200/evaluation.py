"""Full-catalog evaluation front end: score matrices, on-chip top-K, filtered ranks,
and the catalog-sharded multi-GPU form (one all-gather of per-shard top-K keys).

Semantics follow the reference's ranking walk (utils/misc.py:125-146, 213-248) with
ties broken by (score, id); see oracle/kg_oracle.py for the CPU statement.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import functional as KF

KEY_INF = (1 << 64) - 1


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check_catalog(cat):
    if not (cat.is_cuda and cat.dtype == torch.float32 and cat.dim() == 2 and cat.stride(1) == 1):
        raise RuntimeError("kgrec_b200: catalog must be a float32 CUDA matrix with unit column stride")


def keys_to_ids_scores(keys):
    """uint64 keys (stored as int64) -> (ids int64 [nq, k] with -1 for empty, scores float32)."""
    ids = (keys & 0xFFFFFFFF).to(torch.int64)
    bits = ((keys >> 32) & 0xFFFFFFFF).to(torch.int32)
    scores = bits.view(torch.float32)
    empty = keys == -1
    ids = torch.where(empty, torch.full_like(ids, -1), ids)
    scores = torch.where(empty, torch.full_like(scores, float("inf")), scores)
    return ids, scores


def build_filter_csr(query_keys, dicts, device, id_lo=0, id_hi=None):
    """CSR of ids to skip per query: union over `dicts` of dict[key] (the reference's
    `all_dicts`, item_recommendation.py:108-111), restricted to [id_lo, id_hi)."""
    ptr = [0]
    ids = []
    for key in query_keys:
        s = set()
        for d in dicts or ():
            if key in d:
                s.update(d[key])
        row = sorted(i for i in s if i >= id_lo and (id_hi is None or i < id_hi))
        ids.extend(row)
        ptr.append(len(ids))
    return (torch.tensor(ptr, dtype=torch.int64, device=device),
            torch.tensor(ids if ids else [0], dtype=torch.int32, device=device))


def run(T, model, side, q, r, mode, catalog, id_base=0, k=10, filter_csr=None, gumbel_u=None, seed=0,
        qvec=None, gold_scores=None, gold_ids=None, out=None, cat_ids=None):
    """One call into the evaluation kernels.

    mode 'scores' -> [nq, n_cat] float32; 'topk' -> int64-viewed uint64 keys [nq, k];
    'rank' -> int32 counts [nq] (added to `out` when given).
    """
    _check_catalog(catalog)
    lib = _lib.load()
    dev = catalog.device
    nq = (q if q is not None else qvec).shape[0]
    n_cat, cat_ld = catalog.shape[0], catalog.stride(0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ib = (q.element_size() if q is not None else 8)
    if gumbel_u is not None:
        gumbel_u = gumbel_u.to(dev, torch.float32).contiguous()
    if qvec is not None:
        qvec = qvec.to(dev, torch.float32).contiguous()
    if mode == "scores":
        res = torch.empty((nq, n_cat), dtype=torch.float32, device=dev) if out is None else out
        _lib.check(lib.kgrec_eval_scores(C.byref(T), model, side, _ptr(q), _ptr(r), ib, _ptr(qvec), nq,
                                         _ptr(catalog), cat_ld, n_cat, id_base,
                                         _ptr(cat_ids.to(torch.int32).contiguous() if cat_ids is not None else None),
                                         _ptr(gumbel_u), seed, _ptr(res), res.stride(0), stream))
        KF.count_launches(1)
        return res
    if mode == "topk":
        keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
        ws_bytes = lib.kgrec_eval_workspace_bytes(nq, k)
        ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.int64, device=dev)
        fptr, fids = filter_csr if filter_csr is not None else (None, None)
        _lib.check(lib.kgrec_eval_topk(C.byref(T), model, side, _ptr(q), _ptr(r), ib, _ptr(qvec), nq,
                                       _ptr(catalog), cat_ld, n_cat, id_base, k, _ptr(fptr), _ptr(fids),
                                       _ptr(gumbel_u), seed, _ptr(keys), _ptr(ws), ws.numel() * 8, stream))
        KF.count_launches(2)          # scoring kernel + merge of the per-range lists
        return keys
    if mode == "rank":
        counts = torch.zeros(nq, dtype=torch.int32, device=dev) if out is None else out
        _lib.check(lib.kgrec_eval_rank_count(C.byref(T), model, side, _ptr(q), _ptr(r), ib, _ptr(qvec), nq,
                                             _ptr(catalog), cat_ld, n_cat, id_base,
                                             _ptr(gold_scores.contiguous().float()),
                                             _ptr(gold_ids.to(torch.int32).contiguous()), _ptr(counts), stream))
        KF.count_launches(1)
        return counts
    raise ValueError("unknown eval mode %r" % (mode,))


def merge_topk(key_lists):
    """[n_lists, nq, k] keys -> [nq, k] (kgrec_merge_topk)."""
    lib = _lib.load()
    n_lists, nq, k = key_lists.shape
    out = torch.empty((nq, k), dtype=torch.int64, device=key_lists.device)
    _lib.check(lib.kgrec_merge_topk(_ptr(key_lists.contiguous()), n_lists, nq, k, _ptr(out),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    KF.count_launches(1)
    return out


def shard_bounds(n_rows, world, rank):
    """Contiguous row partition: shard g holds ids [g*ceil(n/G), ...)."""
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def sharded_topk(local_keys, group=None):
    """The path's one collective: all-gather the per-shard top-K keys [nq, k] and merge.

    Works on the NCCL backend (CUDA tensors, merge on the GPU) and on gloo (CPU tensors,
    used by the host-logic tests: merge with a stable sort).
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local_keys
    nq, k = local_keys.shape
    gathered = torch.empty((world * nq, k), dtype=local_keys.dtype, device=local_keys.device)
    dist.all_gather_into_tensor(gathered, local_keys.contiguous(), group=group)
    gathered = gathered.view(world, nq, k)
    if local_keys.is_cuda:
        return merge_topk(gathered)
    return merge_topk_host(gathered)


def merge_topk_host(key_lists):
    """Host statement of kgrec_merge_topk (uint64 order on int64 storage)."""
    n_lists, nq, k = key_lists.shape
    flat = key_lists.permute(1, 0, 2).reshape(nq, n_lists * k).numpy().view(np.uint64)
    flat = np.sort(flat, axis=1)[:, :k]
    return torch.from_numpy(flat.view(np.int64).copy())


def sharded_rank_counts(local_counts, group=None):
    """KG mean-rank mode: per-shard counts add (one all-reduce)."""
    import torch.distributed as dist
    if dist.get_world_size(group) > 1:
        dist.all_reduce(local_counts, op=dist.ReduceOp.SUM, group=group)
    return local_counts


# ---- metrics from the reduced outputs (host arithmetic on <= K ids per query) -------------------
def rec_metrics_from_topk(top_ids, gold_sets):
    """P/R/F1/hit/NDCG per user from top-n ids: utils/misc.py:231-248."""
    out = []
    for row, gold in zip(top_ids, gold_sets):
        ids = [int(i) for i in row if i >= 0]
        hits = [1 if i in gold else 0 for i in ids]
        n_hit = sum(hits)
        if n_hit == 0:
            out.append((0.0, 0.0, 0.0, 0, 0.0))
            continue
        p = n_hit / len(hits)
        r = n_hit / len(gold)
        out.append((2 * p * r / (p + r), p, r, 1, _ndcg(hits)))
    return out


def _ndcg(hits):
    """ndcg_at_k(method=0): utils/evaluation.py:41-110."""
    r = np.asarray(hits, dtype=float)

    def dcg(v):
        return v[0] + np.sum(v[1:] / np.log2(np.arange(2, v.size + 1))) if v.size else 0.0
    best = dcg(np.sort(r)[::-1])
    return float(dcg(r) / best) if best else 0.0
