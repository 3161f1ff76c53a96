"""``python-gflags`` as the reference uses it, on top of absl.flags (installed here).

The reference's entry points do ``import gflags``; ``FLAGS = gflags.FLAGS``;
``gflags.DEFINE_*``; ``FLAGS(sys.argv)``; ``FLAGS.FlagValuesDict()`` (jTransUP/models/base.py:22-98,
item_recommendation.py:224).  absl.flags is the same library under its new name, minus the
CamelCase method.  Importing this module also restores ``numpy.asfarray`` (removed in numpy 2;
jTransUP/utils/evaluation.py:69 calls it inside the evaluation worker processes, which are
forked and so inherit the attribute) -- every reference entry point imports gflags first.
"""
import numpy as _np
from absl.flags import *            # noqa: F401,F403
from absl import flags as _flags

FLAGS = _flags.FLAGS
if not hasattr(_flags.FlagValues, "FlagValuesDict"):
    _flags.FlagValues.FlagValuesDict = _flags.FlagValues.flag_values_dict
if not hasattr(_np, "asfarray"):
    _np.asfarray = lambda a, dtype=float: _np.asarray(a, dtype=dtype)   # noqa: E731
