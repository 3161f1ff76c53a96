"""On-disk formats around the hot path.

The reference re-parses its TSV files into Python lists and dicts of sets on every start
(jTransUP/data/load_triple_data.py:5-30, load_rating_data.py:19-38) and feeds the kernels from
those.  Here the same files are parsed once into int32 arrays, cached as one `.npz` next to the
TSV, and exposed in the forms the CUDA path consumes: packed 64-bit keys for the device hash
set (sampling.py), CSR filter lists for the top-K kernel (evaluation.build_filter_csr works on
the dict view, also provided), and the reference's own dict / list views for unchanged code.
Checkpoints use the reference's layout (utils/trainer.py:109-142) so they interchange.
"""
import os

import numpy as np
import torch


def _parse_int_tsv(path, n_cols):
    rows = []
    with open(path, "r", encoding="utf-8") as fin:
        for line in fin:
            parts = line.strip().split("\t")
            if len(parts) != n_cols:          # the reference skips malformed lines the same way
                continue
            rows.append([int(x) for x in parts])
    return np.asarray(rows, dtype=np.int32).reshape(-1, n_cols)


def _cached(path, n_cols, use_cache):
    cache = path + ".kgrec.npz"
    if use_cache and os.path.exists(cache) and os.path.getmtime(cache) >= os.path.getmtime(path):
        with np.load(cache) as z:
            return z["rows"]
    rows = _parse_int_tsv(path, n_cols)
    if use_cache:
        try:
            np.savez(cache, rows=rows)
        except OSError:
            pass                                   # read-only dataset directory: parse every time
    return rows


class TripleFile:
    """`train.dat` / `valid.dat` / `test.dat` of a KG: lines `head \\t tail \\t relation`
    (README.md:45-55; load_triple_data.py:5-30)."""

    def __init__(self, path, use_cache=True):
        self.rows = _cached(path, 3, use_cache)    # [n, 3] int32: h, t, r

    @property
    def total(self):
        return len(self.rows)

    def as_list(self):
        return [tuple(int(v) for v in r) for r in self.rows]

    def head_dict(self):
        """{(t, r): set(heads)} as loadTriples builds it."""
        out = {}
        for h, t, r in self.rows.tolist():
            out.setdefault((t, r), set()).add(h)
        return out

    def tail_dict(self):
        out = {}
        for h, t, r in self.rows.tolist():
            out.setdefault((h, r), set()).add(t)
        return out

    def tensor(self, device="cpu"):
        return torch.from_numpy(self.rows.astype(np.int64)).to(device)


class RatingFile:
    """`train.dat` / eval files of the recommendation task: `user \\t item \\t rating`
    (load_rating_data.py:19-38; the rating value is read and dropped, as in the reference)."""

    def __init__(self, path, use_cache=True):
        self.rows = _cached(path, 3, use_cache)[:, :2].copy()   # [n, 2] int32: u, i

    @property
    def total(self):
        return len(self.rows)

    def as_list(self):
        return [tuple(int(v) for v in r) for r in self.rows]

    def rating_dict(self):
        out = {}
        for u, i in self.rows.tolist():
            out.setdefault(u, set()).add(i)
        return out

    def tensor(self, device="cpu"):
        return torch.from_numpy(self.rows.astype(np.int64)).to(device)


def csr_from_dicts(keys, dicts, device="cpu", id_lo=0, id_hi=None):
    """Vectorised form of evaluation.build_filter_csr for large eval sets."""
    ptr = np.zeros(len(keys) + 1, dtype=np.int64)
    chunks = []
    for n, key in enumerate(keys):
        s = set()
        for d in dicts or ():
            if key in d:
                s.update(d[key])
        row = np.fromiter((i for i in s if i >= id_lo and (id_hi is None or i < id_hi)), dtype=np.int32)
        row.sort()
        chunks.append(row)
        ptr[n + 1] = ptr[n] + len(row)
    ids = np.concatenate(chunks) if chunks and ptr[-1] else np.zeros(1, dtype=np.int32)
    return torch.from_numpy(ptr).to(device), torch.from_numpy(ids).to(device)


def save_checkpoint(path, model, optimizer=None, step=0, best_step=0, best_dev_performance=0.0):
    """The reference's checkpoint dict (utils/trainer.py:115-122), tensors on the CPU."""
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    torch.save({"step": step, "best_step": best_step, "best_dev_performance": best_dev_performance,
                "model_state_dict": sd,
                "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {}}, path)


def load_checkpoint(path, model, optimizer=None):
    """ModelTrainer.load (utils/trainer.py:128-142): strict=False state_dict load; returns the
    bookkeeping fields.  Works on checkpoints written by the reference trainer as well."""
    ck = torch.load(path, map_location="cpu")
    model.load_state_dict(ck["model_state_dict"], strict=False)
    if optimizer is not None and ck.get("optimizer_state_dict"):
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return ck.get("step", 0), ck.get("best_step", 0), ck.get("best_dev_performance", 0.0)
