"""DAGGER learner and training loop with the reference's API (learner/gnn_dagger.py:18-243).

``select_action`` is the inference hot path: for a state that came from the engine-backed env it runs
the sparse CUDA kernels on the history the engine already holds (no dense N x N tensors at all);
otherwise it evaluates ``Actor.forward`` on the dense tensors through the dense CUDA kernel.
``gradient_step`` (training) runs natively (libfgnn.so: fused MLP forward / MSE /
backward kernel + Adam kernel, updating the torch parameters and the torch Adam state in place) whenever every
sampled state carries its aggregated features (states recorded by ``train_dagger`` / ``train_cloning`` on the
engine-backed env); states built from dense tensors fall back to torch autograd on the GPU.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch.optim import Adam

from learner.actor import Actor
from learner.replay_buffer import ReplayBuffer, Transition
from learner.state_with_delay import MultiAgentStateWithDelay
from multiagent_gnn_policies_b200.engine import ActorTrainer


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
        self._trainer = None

    # -- native update ------------------------------------------------------------------------
    def _adam_tensors(self):
        """The torch.optim.Adam state of every parameter (created like Adam.step would on first use), as
        lists in parameter order: the native kernel updates these tensors in place, so ``actor_optim``
        (state_dict, a later torch-side ``step``) stays consistent."""
        params = [p for g in self.actor_optim.param_groups for p in g['params']]
        m, v = [], []
        for p in params:
            st = self.actor_optim.state[p]
            if len(st) == 0:
                st['step'] = torch.tensor(0.0, dtype=torch.float32)
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            m.append(st['exp_avg'])
            v.append(st['exp_avg_sq'])
        return params, m, v

    def _native_gradient_step(self, batch):
        if self._trainer is None:
            self._trainer = ActorTrainer(self.actor.k, self.actor.layers[1], self.actor.n_layers - 1,
                                         device=(torch.device(self.device).index if torch.device(self.device).index is not None
                                                 else torch.cuda.current_device()))
        z = torch.stack([s.aggregated for s in batch.state])                    # (B,K,N,6)
        target = torch.cat(batch.action).to(self.device)                          # (B,1,nA,N)
        params, m, v = self._adam_tensors()
        group = self.actor_optim.param_groups[0]
        step = int(self.actor_optim.state[params[0]]['step'].item()) + 1
        loss, _ = self._trainer.step(z, target, [p.data for p in params], m, v, step=step, lr=group['lr'],
                                     betas=group['betas'], eps=group['eps'])
        for p in params:
            self.actor_optim.state[p]['step'] += 1
        self.actor.native_updates += 1            # the engine-side weight cache keys on this
        return loss.item()

    def _native_supported(self, batch):
        group = self.actor_optim.param_groups
        return (len(group) == 1 and not group[0].get('amsgrad') and not group[0].get('weight_decay')
                and not group[0].get('maximize') and self.actor._engine_supported()
                and torch.device(self.device).type == "cuda"
                and all(getattr(s, "aggregated", None) is not None for s in batch.state))

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
        if self._native_supported(batch):
            return self._native_gradient_step(batch)
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
    raw = getattr(env, 'env', None)
    recording = getattr(raw, 'record_aggregated', False)
    if recording:
        raw.record_aggregated = False            # evaluation states are never stored
    try:
        rewards = _evaluate_episodes(env, learner, args, device, n_episodes)
    finally:
        if recording:
            raw.record_aggregated = True
    return rewards


def _evaluate_episodes(env, learner, args, device, n_episodes):
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

    if hasattr(env, 'env') and hasattr(env.env, 'record_aggregated'):
        env.env.record_aggregated = True          # states carry what the native gradient step consumes

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
