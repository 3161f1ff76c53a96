"""Run the reference's three UNMODIFIED entry points (``baseline/_ref/run_*.py``), either
alone on the host cores or with ``kgrec_b200.dropin`` registering the CUDA model modules
under the reference's module names, and read back what the drivers logged.

Test infrastructure.  ``baseline/_ref`` is produced by ``baseline/make_ref.py`` (a copy of
the reference with the four mechanical torch>=0.4 patches; git-ignored, travels to the GPU
box with the work tree).  Nothing here reads ``/root/reference``.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "joint-kg-recommender_b200")
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import make_ref          # noqa: E402
import synth_dataset     # noqa: E402

DRIVERS = {"kg": "run_knowledge_representation.py", "rec": "run_item_recommendation.py",
           "joint": "run_knowledgable_recommendation.py"}

_NUM = r"([-+0-9.eE]+|nan)"
_PAT = {
    "train_loss": re.compile(r"INFO - train loss:%s!" % _NUM),
    "joint_loss": re.compile(r"rec train loss:%s, kg train loss:%s!" % (_NUM, _NUM)),
    "kg": re.compile(r"avg hit:%s, avg mean rank:%s, topn" % (_NUM, _NUM)),
    "kg_head": re.compile(r"head hit:%s, head mean rank:%s, topn" % (_NUM, _NUM)),
    "kg_tail": re.compile(r"tail hit:%s, tail mean rank:%s, topn" % (_NUM, _NUM)),
    "rec": re.compile(r"f1:%s, p:%s, r:%s, hit:%s, ndcg:%s, topn" % ((_NUM,) * 5)),
}


def reference_available():
    return make_ref.available()


def parse_log(path):
    out = {k: [] for k in _PAT}
    with open(path, encoding="utf-8", errors="replace") as f:
        for line in f:
            for k, pat in _PAT.items():
                m = pat.search(line)
                if m:
                    out[k].append(tuple(float(x) for x in m.groups()))
    return out


def run_driver(kind, flags, log_dir, name, dropin=False, cpu=False, timeout=1500):
    """Run one entry point to completion; returns (parsed log, checkpoint path, log path)."""
    os.makedirs(log_dir, exist_ok=True)
    env = dict(os.environ)
    paths = make_ref.env_paths() + ([PKG] if dropin else [])
    env["PYTHONPATH"] = os.pathsep.join(paths + [env.get("PYTHONPATH", "")])
    env["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"     # the trainer's checkpoints hold numpy scalars (trainer.py:115-122)
    env["PYTHONWARNINGS"] = "ignore"
    if cpu:
        env["CUDA_VISIBLE_DEVICES"] = ""              # USE_CUDA is latched at import (utils/misc.py:11)
    script = os.path.join(make_ref.DEST, DRIVERS[kind])
    cmd = [sys.executable] + (["-m", "kgrec_b200.dropin"] if dropin else []) + [script]
    cmd += list(flags) + ["-log_path", log_dir + os.sep, "-experiment_name", name, "-nohas_visualization"]
    log = os.path.join(log_dir, name + ".log")
    if os.path.exists(log):
        os.remove(log)
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("%s exited %d\n--- stderr tail ---\n%s" % (" ".join(cmd), r.returncode, r.stderr[-4000:]))
    return parse_log(log), os.path.join(log_dir, name + ".ckpt"), log


def make_dataset(root, **kw):
    return synth_dataset.write_dataset(root, "ml1m", **kw)
