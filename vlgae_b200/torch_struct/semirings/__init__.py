from .semirings import NEGINF, LogSemiring, MaxSemiring, Semiring  # noqa: F401
