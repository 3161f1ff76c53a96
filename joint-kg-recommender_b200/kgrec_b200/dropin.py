"""Make the reference's drivers pick up the CUDA engine without editing them.

``jTransUP/models/base.py:10-16`` imports ``jTransUP.models.{transE,transH,transR,transUP,
jTransUP}`` by name and ``init_model`` (base.py:128-166) calls ``<module>.build_model``.
``install()`` registers this package's modules under those names in ``sys.modules`` BEFORE
``jTransUP.models.base`` is imported, so ``run_item_recommendation.py``,
``run_knowledge_representation.py`` and ``run_knowledgable_recommendation.py`` and the
``-model_type`` dispatch work unchanged:

    python -m kgrec_b200.dropin /path/to/reference/run_knowledge_representation.py -model_type transe ...

When the reference package itself is importable it keeps every other module (data loaders,
trainer, drivers); only the five model modules are replaced.
"""
import importlib
import runpy
import sys
import types

_NAMES = ("transE", "transH", "transR", "transUP", "jTransUP")


def install():
    """Register the replacements; returns what was shadowed (for uninstall)."""
    saved = {}
    for pkg in ("jTransUP", "jTransUP.models"):
        if pkg not in sys.modules:
            try:
                importlib.import_module(pkg)          # the real reference package, if on sys.path
            except ImportError:
                mod = types.ModuleType(pkg)           # a bare namespace otherwise
                mod.__path__ = []
                saved[pkg] = None
                sys.modules[pkg] = mod
    for name in _NAMES:
        full = "jTransUP.models." + name
        saved[full] = sys.modules.get(full)
        mod = importlib.import_module("kgrec_b200.models." + name)
        sys.modules[full] = mod
        setattr(sys.modules["jTransUP.models"], name, mod)
    return saved


def uninstall(saved):
    """Undo install(): sys.modules entries AND the attributes install() set on the `jTransUP.models` package
    (`from jTransUP.models import transE` reads the attribute before sys.modules)."""
    for full, old in saved.items():
        pkg_name, _, leaf = full.rpartition(".")
        pkg = sys.modules.get(pkg_name)
        if old is None:
            sys.modules.pop(full, None)
            if pkg is not None and leaf in _NAMES and hasattr(pkg, leaf):
                delattr(pkg, leaf)
        else:
            sys.modules[full] = old
            if pkg is not None and leaf in _NAMES:
                setattr(pkg, leaf, old)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m kgrec_b200.dropin <reference run_*.py> [flags...]")
    install()
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
