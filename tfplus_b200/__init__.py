"""tfplus_b200 — B200-native KvVariable sparse-embedding hot path.

HBM-resident open-addressing table + hand-written sm_100a kernels behind
TFPlus's own op surface.  `ops` mirrors gen_kv_variable_ops; the KvVariable /
get_kv_variable / embedding_lookup / optimizer layer mirrors
tfplus/kv_variable/python/{ops,training}.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
from .kv_variable import (IndexedSlices, KvVariable, PartitionedKvVariable,  # noqa: F401,E402
                          constant_initializer, fixed_size_partitioner, get_kv_variable,
                          ones_initializer, random_normal_initializer, reset_kv_variable_store,
                          set_training, zeros_initializer)
from .embedding_ops import (embedding_lookup, embedding_lookup_sparse,  # noqa: F401,E402
                            safe_embedding_lookup_sparse)
from . import training  # noqa: F401,E402
