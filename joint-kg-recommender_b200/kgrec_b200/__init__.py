"""kgrec_b200 -- B200-native scoring engine behind the model API of
TaoMiner/joint-kg-recommender (TransE / TransH / TransR / TUP / KTUP).

Python here is the host-side mirror of the reference's nn.Module protocol; all
arithmetic happens in the CUDA library declared in include/kgrec_b200.h.
"""
from . import _lib  # noqa: F401
from .models import transE, transH, transR, transUP, jTransUP  # noqa: F401
from .models.transE import TransEModel
from .models.transH import TransHModel
from .models.transR import TransRModel
from .models.transUP import TransUPModel
from .models.jTransUP import jTransUPModel

__all__ = ["TransEModel", "TransHModel", "TransRModel", "TransUPModel", "jTransUPModel"]
