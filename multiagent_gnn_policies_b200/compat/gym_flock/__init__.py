"""``gym_flock`` call surface on the CUDA engine (the real package is an un-vendored dependency of the
reference, README.md:7).  FlockingRelative-v0 is the hot path; Leader / TwoFlocks /
Stochastic are the variants the other cfgs name (8f row f3).  The AirSim ids need an external simulator and
stay out of scope."""
import gym

from gym_flock import envs
from gym_flock.envs import FlockingRelativeEnv, FlockingLeaderEnv, FlockingTwoFlocksEnv, FlockingStochasticEnv

gym.register("FlockingRelative-v0", FlockingRelativeEnv, max_episode_steps=200)
gym.register("FlockingLeader-v0", FlockingLeaderEnv, max_episode_steps=200)
gym.register("FlockingTwoFlocks-v0", FlockingTwoFlocksEnv, max_episode_steps=200)
gym.register("FlockingStochastic-v0", FlockingStochasticEnv, max_episode_steps=200)
