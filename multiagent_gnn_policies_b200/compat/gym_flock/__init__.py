"""``gym_flock`` call surface on the CUDA engine (the real package is an un-vendored dependency of the
reference, README.md:7).  Only FlockingRelative-v0 is in scope."""
import gym

from gym_flock import envs
from gym_flock.envs import FlockingRelativeEnv

gym.register("FlockingRelative-v0", FlockingRelativeEnv, max_episode_steps=200)
