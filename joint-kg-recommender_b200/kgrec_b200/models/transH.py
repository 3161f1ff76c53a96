"""TransH on the CUDA engine.  Mirrors jTransUP/models/transH.py (constructor 17-56,
forward 58-71, evaluateHead/Tail 73-121; projection utils/misc.py:18-19)."""
from .. import _lib
from .transE import KGModelBase


def build_model(FLAGS, user_total, item_total, entity_total, relation_total, i_map=None, e_map=None, new_map=None):
    return TransHModel(L1_flag=FLAGS.L1_flag, embedding_size=FLAGS.embedding_size,
                       ent_total=entity_total, rel_total=relation_total)


class TransHModel(KGModelBase):
    MODEL = _lib.TRANSH
    TABLES = {"ent": "ent_embeddings", "rel": "rel_embeddings", "norm": "norm_embeddings"}

    def __init__(self, L1_flag, embedding_size, ent_total, rel_total):
        super().__init__(L1_flag, embedding_size, ent_total, rel_total)
        self._finish_init()

    def _table_specs(self):
        # per-relation hyperplane normals, unit length at init, never re-normalised in forward
        return super()._table_specs() + [("norm_embeddings", self.rel_total, self.embedding_size, True)]
