#!/usr/bin/env python
"""Synthetic datasets in the reference's on-disk layout (README.md:45-72 of the reference).

The reference ships no data (its README links an archive that cannot be fetched here), so
every run of its drivers -- alone or with the CUDA modules swapped in -- reads a dataset
written by this module:

    <root>/<name>/train.dat valid.dat test.dat      "user \t item \t rating"       (load_rating_data.py:19-38)
    <root>/<name>/u_map.dat i_map.dat               "mapped_id \t original_id"     (load_rating_data.py:6-16)
    <root>/<name>/i2kg_map.tsv                      "orig_item \t title \t uri"    (load_kg_rating_data.py:5-18)
    <root>/<name>/kg/train.dat valid.dat test.dat   "head \t tail \t relation"     (load_triple_data.py:5-30)
    <root>/<name>/kg/e_map.dat r_map.dat            "mapped_id \t uri"             (load_triple_data.py:32-43)

The data has planted structure (users and items in taste clusters; relations shift an entity's
cluster along a line of clusters) so that training moves the metrics away from chance, which makes
a reference-vs-CUDA comparison of logged metrics meaningful.  Splits are 70 / 10 / 20.
``name`` must be one of the reference's ``-dataset`` enum values (base.py:26); "ml1m" is used.
"""
import os

import numpy as np


def _unique_pairs(a, b, nb):
    key = np.unique(a.astype(np.int64) * nb + b)
    return (key // nb).astype(np.int64), (key % nb).astype(np.int64)


def _split(n, rng):
    perm = rng.permutation(n)
    a, b = int(0.7 * n), int(0.8 * n)
    return perm[:a], perm[a:b], perm[b:]


def _write_rows(path, cols):
    arr = np.stack(cols, axis=1)
    np.savetxt(path, arr, fmt="%d", delimiter="\t")


def write_dataset(root, name="ml1m", users=300, items=400, ratings=8000, entities=600, relations=8,
                  triples=6000, clusters=8, aligned_frac=0.7, seed=0, kg=True, rec=True):
    """Write the files; returns a dict of the sizes actually written."""
    rng = np.random.RandomState(seed)
    base = os.path.join(root, name)
    os.makedirs(os.path.join(base, "kg"), exist_ok=True)
    info = {"path": base}
    if rec:
        u = rng.randint(0, users, ratings)
        in_cluster = rng.rand(ratings) < 0.8
        per = max(1, items // clusters)
        i_c = (u % clusters) * per + rng.randint(0, per, ratings)
        i = np.where(in_cluster, np.minimum(i_c, items - 1), rng.randint(0, items, ratings))
        u, i = _unique_pairs(u, i, items)
        # every user and item id must appear in the maps; ids are dense 0..n-1
        tr, va, te = _split(u.size, rng)
        for fname, sel in (("train.dat", tr), ("valid.dat", va), ("test.dat", te)):
            _write_rows(os.path.join(base, fname), [u[sel], i[sel], np.ones(sel.size, np.int64)])
        with open(os.path.join(base, "u_map.dat"), "w") as f:
            f.write("".join("%d\tu%d\n" % (k, k) for k in range(users)))
        with open(os.path.join(base, "i_map.dat"), "w") as f:
            f.write("".join("%d\ti%d\n" % (k, k) for k in range(items)))
        info.update(users=users, items=items, ratings=int(u.size), rating_train=int(tr.size))
    if kg:
        h = rng.randint(0, entities, triples)
        r = rng.randint(0, relations, triples)
        per = max(1, entities // clusters)
        # relation r shifts an entity's cluster by r - relations // 2 along a line of clusters (additive, so
        # translation models can represent it); shifts that leave the line fall back to a random tail
        c_t = np.minimum(h // per, clusters - 1) + (r - relations // 2)
        ok = (c_t >= 0) & (c_t < clusters) & (rng.rand(triples) < 0.9)
        t_c = np.clip(c_t, 0, clusters - 1) * per + rng.randint(0, per, triples)
        t = np.where(ok, np.minimum(t_c, entities - 1), rng.randint(0, entities, triples))
        keep = h != t
        h, r, t = h[keep], r[keep], t[keep]
        key = np.unique((h.astype(np.int64) * relations + r) * entities + t)
        h, r, t = key // (relations * entities), (key // entities) % relations, key % entities
        tr, va, te = _split(h.size, rng)
        for fname, sel in (("train.dat", tr), ("valid.dat", va), ("test.dat", te)):
            _write_rows(os.path.join(base, "kg", fname), [h[sel], t[sel], r[sel]])
        with open(os.path.join(base, "kg", "e_map.dat"), "w") as f:
            f.write("".join("%d\thttp://kg/e%d\n" % (k, k) for k in range(entities)))
        with open(os.path.join(base, "kg", "r_map.dat"), "w") as f:
            f.write("".join("%d\thttp://kg/r%d\n" % (k, k) for k in range(relations)))
        info.update(entities=entities, relations=relations, triples=int(h.size), triple_train=int(tr.size))
    if kg and rec:
        n_al = min(int(aligned_frac * items), entities)
        al_items = rng.permutation(items)[:n_al]
        al_ents = rng.permutation(entities)[:n_al]
        with open(os.path.join(base, "i2kg_map.tsv"), "w") as f:
            f.write("".join("i%d\ttitle %d\thttp://kg/e%d\n" % (a, a, b) for a, b in zip(al_items, al_ents)))
        info.update(aligned=int(n_al))
    return info


if __name__ == "__main__":
    import sys
    print(write_dataset(sys.argv[1] if len(sys.argv) > 1 else "/tmp/kgrec_synth"))
