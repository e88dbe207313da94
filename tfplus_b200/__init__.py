"""tfplus_b200 — B200-native KvVariable sparse-embedding hot path.

HBM-resident open-addressing table + hand-written sm_100a kernels behind
TFPlus's own op surface.  `ops` mirrors gen_kv_variable_ops; the KvVariable /
get_kv_variable / embedding_lookup / optimizer layer mirrors
tfplus/kv_variable/python/{ops,training}.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
