"""DAGGER learner and training loop with the reference's API (learner/gnn_dagger.py:18-243).

``select_action`` is the inference hot path: for a state that came from the engine-backed env it runs
the sparse CUDA kernels on the history the engine already holds (no dense N x N tensors at all);
otherwise it evaluates ``Actor.forward`` on the dense tensors through the dense CUDA kernel.
``gradient_step`` (training, SURVEY.md 8f "next" row) uses torch autograd on the GPU.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch.optim import Adam

from learner.actor import Actor
from learner.replay_buffer import ReplayBuffer, Transition
from learner.state_with_delay import MultiAgentStateWithDelay


class DAGGER(object):

    def __init__(self, device, args, k=None):
        n_s = args.getint('n_states')
        n_a = args.getint('n_actions')
        k = k or args.getint('k')
        hidden_size = args.getint('hidden_size')
        n_layers = args.getint('n_layers') or 2
        self.n_agents = args.getint('n_agents')
        self.n_states = n_s
        self.n_actions = n_a
        self.device = device
        self.actor = Actor(n_s, n_a, [hidden_size] * n_layers, k, 0).to(self.device)
        self.actor_optim = Adam(self.actor.parameters(), lr=args.getfloat('actor_lr'))
        self.gamma = args.getfloat('gamma')
        self.tau = args.getfloat('tau')

    def select_action(self, state):
        """(N, n_actions) tensor on ``self.device`` -- callers do ``.cpu().numpy()`` (gnn_dagger.py:161)."""
        self.actor.eval()
        engine = getattr(state, "engine", None)
        with torch.no_grad():
            if (engine is not None and engine.step_index == state.step and engine.k == self.actor.k
                    and engine.hidden == self.actor.layers[1] and engine.n_layers == self.actor.n_layers - 1):
                self.actor.sync_engine(engine)
                mu = engine.policy().view(self.n_agents, self.n_actions)
            else:
                mu = self.actor(state.delay_state, state.delay_gso)         # (B,1,nA,N)
                mu = mu.permute(0, 1, 3, 2).reshape(self.n_agents, self.n_actions)
        self.actor.train()
        return mu.data

    def gradient_step(self, batch):
        delay_gso_batch = torch.cat(tuple(s.delay_gso for s in batch.state)).to(self.device)
        delay_state_batch = torch.cat(tuple(s.delay_state for s in batch.state)).to(self.device)
        actor_batch = self.actor(delay_state_batch, delay_gso_batch)
        optimal_action_batch = torch.cat(batch.action).to(self.device)
        self.actor_optim.zero_grad()
        policy_loss = F.mse_loss(actor_batch, optimal_action_batch)
        policy_loss.backward()
        self.actor_optim.step()
        return policy_loss.item()

    def save_model(self, env_name, suffix="", actor_path=None):
        os.makedirs('models/', exist_ok=True)
        if actor_path is None:
            actor_path = "models/actor_{}_{}".format(env_name, suffix)
        print('Saving model to {}'.format(actor_path))
        torch.save(self.actor.state_dict(), actor_path)

    def load_model(self, actor_path, map_location):
        if actor_path is not None:
            self.actor.load_state_dict(torch.load(actor_path, map_location))
            self.actor.to(self.device)


def _evaluate(env, learner, args, device, n_episodes):
    """Learner-only rollouts (gnn_dagger.py:190-231): list of episode rewards."""
    rewards = []
    for _ in range(n_episodes):
        total = 0
        state = MultiAgentStateWithDelay(device, args, env.reset(), prev_state=None)
        done = False
        while not done:
            action = learner.select_action(state)
            next_state, reward, done, _ = env.step(action.cpu().numpy())
            state = MultiAgentStateWithDelay(device, args, next_state, prev_state=state)
            total += reward
        rewards.append(total)
    return rewards


def train_dagger(env, args, device):
    debug = args.getboolean('debug')
    memory = ReplayBuffer(max_size=args.getint('buffer_size'))
    learner = DAGGER(device, args)
    n_a = args.getint('n_actions')
    n_agents = args.getint('n_agents')
    batch_size = args.getint('batch_size')
    n_train_episodes = args.getint('n_train_episodes')
    beta_coeff = args.getfloat('beta_coeff')
    test_interval = args.getint('test_interval')
    n_test_episodes = args.getint('n_test_episodes')

    total_numsteps, updates, beta = 0, 0, 1
    stats = {'mean': -1.0 * np.inf, 'std': 0}
    for episode in range(n_train_episodes):
        beta = max(beta * beta_coeff, 0.5)
        state = MultiAgentStateWithDelay(device, args, env.reset(), prev_state=None)
        done = False
        policy_loss_sum = 0
        while not done:
            optimal_action = env.env.controller()
            if np.random.binomial(1, beta) > 0:
                action = optimal_action                       # expert drives with probability beta
            else:
                action = learner.select_action(state).cpu().numpy()
            next_state, reward, done, _ = env.step(action)
            next_state = MultiAgentStateWithDelay(device, args, next_state, prev_state=state)
            total_numsteps += 1
            notdone = torch.Tensor([not done]).to(device)
            reward = torch.Tensor([reward]).to(device)
            # the expert label, (N,nA) -> (1,1,nA,N)
            label = torch.Tensor(optimal_action).to(device).transpose(1, 0).reshape((1, 1, n_a, n_agents))
            memory.insert(Transition(state, label, notdone, next_state, reward))
            state = next_state
        if memory.curr_size > batch_size:
            for _ in range(args.getint('updates_per_step')):
                batch = Transition(*zip(*memory.sample(batch_size)))
                policy_loss_sum += learner.gradient_step(batch)
                updates += 1
        if episode % test_interval == 0 and debug:
            mean_reward = np.mean(_evaluate(env, learner, args, device, n_test_episodes))
            print("Episode: {}, updates: {}, total numsteps: {}, reward: {}, policy loss: {}".format(
                episode, updates, total_numsteps, mean_reward, policy_loss_sum))
    test_rewards = _evaluate(env, learner, args, device, n_test_episodes)
    stats['mean'] = np.mean(test_rewards)
    stats['std'] = np.std(test_rewards)
    if debug and args.get('fname'):
        learner.save_model(args.get('env'), suffix=args.get('fname'))
    env.close()
    return stats
