"""synthetic_state for scripts that must not import oracle/: re-exported through tests/ (test infrastructure)."""
from oracle.flock_env import synthetic_state  # noqa: F401
