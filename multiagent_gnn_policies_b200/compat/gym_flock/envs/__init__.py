from gym_flock.envs.flocking_relative import FlockingRelativeEnv, EngineState
from gym_flock.envs.flocking_variants import FlockingLeaderEnv, FlockingTwoFlocksEnv, FlockingStochasticEnv

__all__ = ["FlockingRelativeEnv", "FlockingLeaderEnv", "FlockingTwoFlocksEnv", "FlockingStochasticEnv", "EngineState"]
