"""The four TFPlus optimizers in scope, for KvVariables on the device
(mirror of tfplus/kv_variable/python/training/__init__.py:17-22)."""
from .optimizers import (AdagradOptimizer, AdamOptimizer, GradientDescentOptimizer,  # noqa: F401
                         GroupAdamOptimizer, SparseGroupFtrlOptimizer)
