from .sampling import CFGSampler  # noqa: F401
