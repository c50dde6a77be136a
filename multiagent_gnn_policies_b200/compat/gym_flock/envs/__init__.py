from gym_flock.envs.flocking_relative import FlockingRelativeEnv, EngineState

__all__ = ["FlockingRelativeEnv", "EngineState"]
