#!/usr/bin/env python
"""Build ``baseline/_ref/``: the reference itself, runnable on today's torch / numpy.

    python baseline/make_ref.py [--src /root/reference] [--force]

The reference (TaoMiner/joint-kg-recommender) is pure Python written for PyTorch 0.3.  This
script COPIES its ``run_*.py`` entry points and the ``jTransUP`` package from the read-only
checkout into the git-ignored ``baseline/_ref/`` (nothing of it enters this repository's
history; the directory travels to the GPU box with the work tree the same way the built
``.so`` does) and applies the four mechanical patches SURVEY.md section 8c lists -- the only
edits needed for its three drivers to run on torch >= 0.4:

  1. ``losses.data[0]`` -> ``losses.item()``      (0-d tensors cannot be indexed any more)
       item_recommendation.py:193, knowledge_representation.py:217,
       knowledgable_recommendation.py:405,407
  2. ``self.paddingItems(<ids>.data, ...)`` -> ``<ids>.data.tolist()``  (iterating a tensor
       yields 0-d tensors, which miss the ``i_map`` dict)  jTransUP.py:127,174,323; CKE.py:114

Everything else the reference needs and this image lacks is supplied from OUTSIDE its tree
by ``baseline/shims/`` (``gflags`` -> absl.flags, a ``visdom`` stub, ``numpy.asfarray``).

What it is used for: (a) the ``--impl reference`` arm and ``cpu_baseline`` of bench.py time
these unmodified classes on the host cores; (b) tests run the three unmodified drivers twice
-- alone on the CPU, and with ``kgrec_b200.dropin`` swapping in the CUDA modules -- and compare.
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SHIMS = os.path.join(HERE, "shims")
STAMP = os.path.join(DEST, ".kgrec_ref_stamp")

PATCHES = {
    "jTransUP/models/item_recommendation.py": [(r"losses\.data\[0\]", "losses.item()", 1)],
    "jTransUP/models/knowledge_representation.py": [(r"losses\.data\[0\]", "losses.item()", 1)],
    "jTransUP/models/knowledgable_recommendation.py": [(r"losses\.data\[0\]", "losses.item()", 2)],
    "jTransUP/models/jTransUP.py": [
        (r"self\.paddingItems\(i_ids\.data,", "self.paddingItems(i_ids.data.tolist(),", 2),
        (r"self\.paddingItems\(all_i_ids\.data if", "self.paddingItems(all_i_ids.data.tolist() if", 1)],
    "jTransUP/models/CKE.py": [(r"self\.paddingItems\(i_ids\.data,", "self.paddingItems(i_ids.data.tolist(),", 1)],
}


def available():
    return os.path.exists(STAMP)


def env_paths():
    """sys.path / PYTHONPATH entries that make ``baseline/_ref`` importable: shims first."""
    return [SHIMS, DEST]


def make(src="/root/reference", force=False, verbose=True):
    if available() and not force:
        return DEST
    if not os.path.isdir(os.path.join(src, "jTransUP")):
        raise FileNotFoundError("reference checkout not found at %s" % src)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    keep = lambda d, names: [n for n in names if n == "__pycache__" or n.startswith(".")          # noqa: E731
                             or (os.path.isfile(os.path.join(d, n)) and not n.endswith(".py"))]
    shutil.copytree(os.path.join(src, "jTransUP"), os.path.join(DEST, "jTransUP"), ignore=keep)
    for name in sorted(os.listdir(src)):
        if name.startswith("run_") and name.endswith(".py"):
            shutil.copyfile(os.path.join(src, name), os.path.join(DEST, name))
    for root, _dirs, files in os.walk(DEST):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    for rel, subs in PATCHES.items():
        path = os.path.join(DEST, rel)
        text = open(path, encoding="utf-8").read()
        for pat, repl, count in subs:
            text, n = re.subn(pat, repl, text)
            if n != count:
                raise RuntimeError("%s: expected %d site(s) of /%s/, found %d" % (rel, count, pat, n))
        with open(path, "w", encoding="utf-8") as f:
            f.write(text)
    with open(STAMP, "w") as f:
        f.write("source=%s\npatches=%d files\n" % (src, len(PATCHES)))
    if verbose:
        print("[make_ref] %s <- %s (%d patched files)" % (DEST, src, len(PATCHES)), flush=True)
    return DEST


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "/root/reference"
    make(src, force="--force" in sys.argv)
