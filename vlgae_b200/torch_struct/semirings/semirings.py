"""The two semirings the hot path reaches (reference: semirings/semirings.py:127-148,173-207).

Only their identity matters here -- the arithmetic runs inside the CUDA kernels -- but the module
keeps the reference's writable global ``NEGINF`` because ``src.setup_inf`` rebinds it
(/root/reference/src/__init__.py:113-120), and the class attribute ``zero`` that stays frozen at the
import-time value (it is what the single-root mask writes, dmv.py:63).
"""
NEGINF = -1e12


class Semiring:
    zero = None
    name = "abstract"


class LogSemiring(Semiring):
    """(logsumexp, +): gradients are marginals."""
    zero = NEGINF
    name = "log"


class MaxSemiring(Semiring):
    """(max, +) with torch.max's first-index tie rule: gradients are the argmax indicator."""
    zero = NEGINF
    name = "max"
