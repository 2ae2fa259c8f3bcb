"""Mirror of the reference's ``src/model/torch_struct`` package surface that ``src/model`` imports
(/root/reference/src/model/ldndmv.py:21-22, joint.py:20, model/dmv.py:16, src/__init__.py:113-117).

Same names, shapes, dtypes and error behaviour; the chart DP underneath is the sm_100a kernel set in
``vlgae_b200/csrc`` instead of O(N) ATen launches plus autograd through the chart.
"""
from . import semirings  # noqa: F401  (src.setup_inf rebinds semirings.semirings.NEGINF)
from .distributions import DMV1o, DependencyCRF, StructDistribution
from .semirings import LogSemiring, MaxSemiring

version = "0.4"  # the reference vendors pytorch-struct 0.4 (torch_struct/__init__.py:20)

__all__ = ["DMV1o", "DependencyCRF", "StructDistribution", "LogSemiring", "MaxSemiring", "semirings", "version"]
