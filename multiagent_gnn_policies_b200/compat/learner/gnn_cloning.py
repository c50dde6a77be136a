"""Behaviour cloning (reference: learner/gnn_cloning.py:17-213): the learner class is DAGGER's; the
rollout always steps the expert action and the best evaluation so far is kept."""
import numpy as np
import torch

from learner.gnn_dagger import DAGGER, _evaluate
from learner.replay_buffer import ReplayBuffer, Transition
from learner.state_with_delay import MultiAgentStateWithDelay


class ImitationLearning(DAGGER):
    """learner/gnn_cloning.py:17-51: ``(device, args)`` only, and always TWO hidden layers of ``hidden_size``
    (``hidden_layers = [hidden_size, hidden_size]``, gnn_cloning.py:40) whatever the cfg's ``n_layers`` says; no
    ``k=`` override.  Everything else (select_action / gradient_step / save / load) is DAGGER's."""

    def __init__(self, device, args):
        class _TwoLayers:                       # the cfg section with n_layers pinned to 2
            def __init__(self, inner):
                self._inner = inner

            def getint(self, key, *a, **kw):
                return 2 if key == 'n_layers' else self._inner.getint(key, *a, **kw)

            def __getattr__(self, name):
                return getattr(self._inner, name)

        super().__init__(device, _TwoLayers(args))


def train_cloning(env, args, device):
    debug = args.getboolean('debug')
    memory = ReplayBuffer(max_size=args.getint('buffer_size'))
    learner = ImitationLearning(device, args)
    n_a = args.getint('n_actions')
    n_agents = args.getint('n_agents')
    batch_size = args.getint('batch_size')
    updates_per_step = args.getint('updates_per_step')
    n_train_episodes = args.getint('n_train_episodes')
    test_interval = args.getint('test_interval')
    n_test_episodes = args.getint('n_test_episodes')

    if hasattr(env, 'env') and hasattr(env.env, 'record_aggregated'):
        env.env.record_aggregated = True          # states carry what the native gradient step consumes

    total_numsteps, updates = 0, 0
    stats = {'mean': -1.0 * np.inf, 'std': 0}
    for episode in range(n_train_episodes):
        state = MultiAgentStateWithDelay(device, args, env.reset(), prev_state=None)
        done = False
        policy_loss_sum = 0
        while not done:
            optimal_action = env.env.controller()
            next_state, reward, done, _ = env.step(optimal_action)
            next_state = MultiAgentStateWithDelay(device, args, next_state, prev_state=state)
            total_numsteps += 1
            notdone = torch.Tensor([not done]).to(device)
            reward = torch.Tensor([reward]).to(device)
            label = torch.Tensor(optimal_action).to(device).transpose(1, 0).reshape((1, 1, n_a, n_agents))
            memory.insert(Transition(state, label, notdone, next_state, reward))
            state = next_state
        if memory.curr_size > batch_size:
            for _ in range(updates_per_step):
                batch = Transition(*zip(*memory.sample(batch_size)))
                policy_loss_sum += learner.gradient_step(batch)
                updates += 1
        if episode % test_interval == 0:
            rewards = _evaluate(env, learner, args, device, n_test_episodes)
            mean_reward = np.mean(rewards)
            if stats['mean'] < mean_reward:
                stats['mean'] = mean_reward
                stats['std'] = np.std(rewards)
                if debug and args.get('fname'):
                    learner.save_model(args.get('env'), suffix=args.get('fname'))
            if debug:
                print("Episode: {}, updates: {}, total numsteps: {}, reward: {}, policy loss: {}".format(
                    episode, updates, total_numsteps, mean_reward, policy_loss_sum))
    env.close()
    return stats
