"""Expert-controller evaluation (reference: learner/gnn_baseline.py:4-27)."""
import numpy as np


def train_baseline(env, args):
    n_test_episodes = args.getint('n_test_episodes')
    centralized = args.getboolean('centralized')
    rewards = []
    for _ in range(n_test_episodes):
        total = 0
        env.reset()
        done = False
        while not done:
            _, reward, done, _ = env.step(env.env.controller(centralized))
            total += reward
        rewards.append(total)
    env.close()
    return {'mean': np.mean(rewards), 'std': np.std(rewards)}
