"""``learner`` package of the reference (learner/*.py) on the CUDA engine."""
