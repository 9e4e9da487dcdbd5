"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the hot path.  The product package
(`unlearn_ft_b200`) never imports this; only `tests/`, `__graft_entry__.smoke()` and the cpu_baseline / reference
arm of `bench.py` do, and only as the checker.
"""
