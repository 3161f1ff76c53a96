"""TransH on the CUDA engine.  Mirrors jTransUP/models/transH.py (constructor 17-56,
forward 58-71, evaluateHead/Tail 73-121; projection utils/misc.py:18-19)."""
from .. import _lib
from .base import _embedding, _init_table
from .transE import KGModelBase


def build_model(FLAGS, user_total, item_total, entity_total, relation_total, i_map=None, e_map=None, new_map=None):
    return TransHModel(L1_flag=FLAGS.L1_flag, embedding_size=FLAGS.embedding_size,
                       ent_total=entity_total, rel_total=relation_total)


class TransHModel(KGModelBase):
    MODEL = _lib.TRANSH
    TABLES = {"ent": "ent_embeddings", "rel": "rel_embeddings", "norm": "norm_embeddings"}

    def __init__(self, L1_flag, embedding_size, ent_total, rel_total):
        super().__init__(L1_flag, embedding_size, ent_total, rel_total)
        # per-relation hyperplane normals, unit length at init, never re-normalised in forward
        self.norm_embeddings = _embedding(_init_table(rel_total, embedding_size))
        self._finish_init()
