"""DAGGER for B parallel episodes, entirely on the device (BASELINE config C3: 256 episodes x N = 1k agents).

The reference trains one episode at a time from the host (learner/gnn_dagger.py:126-243): expert action
(:156), beta-mixed rollout (:158-163), transition into a replay buffer of dense states (:170-178), then
``updates_per_step`` gradient steps per episode (:183-188).  ``DeviceDagger`` runs the same algorithm for
B block-diagonal episodes on one engine handle with nothing crossing PCIe inside an episode:

* expert labels      -> ``fgnn_controller``       (k_controller, decentralised or centralised)
* learner actions    -> ``fgnn_policy``           (hops + readout), which also leaves the aggregated features
* stored transition  -> the aggregated features z (K, B*N, 6) + the label (B*N, 2): a device ring buffer
* beta mixing        -> per EPISODE Bernoulli(beta) mask, ``torch.where`` on device tensors (plumbing)
* env.step           -> ``fgnn_integrate`` + ``fgnn_build_graph``
* gradient_step      -> ``fgnn_trainer_step`` on ``batch_size`` sampled (step, episode) states

PyTorch is plumbing here (device buffers, RNG for the sampling, the Adam state tensors).
"""
import numpy as np

from multiagent_gnn_policies_b200.engine import ActorTrainer, FlockEngine


class DeviceDagger:
    def __init__(self, n_agents, n_episodes, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, lr=5e-5,
                 buffer_steps=200, batch_size=20, beta_coeff=0.993, centralized=False, max_accel=1.0, device=0,
                 seed=11, edge_capacity=0):
        import torch
        self.torch = torch
        self.n_agents, self.n_episodes, self.k = n_agents, n_episodes, k
        self.hidden, self.n_layers = hidden, n_layers
        self.device = torch.device("cuda", device)
        self.engine = FlockEngine(n_agents=n_agents, n_episodes=n_episodes, k=k, hidden=hidden, n_layers=n_layers,
                                  comm_radius=comm_radius, dt=dt, device=device, edge_capacity=edge_capacity)
        self.trainer = ActorTrainer(k, hidden, n_layers, device=device)
        self.lr, self.batch_size, self.beta_coeff = lr, batch_size, beta_coeff
        self.centralized, self.max_accel = centralized, max_accel
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        # parameters in the reference's conv layout + Adam state, all on the device
        torch.manual_seed(seed)
        dims = [6] + [hidden] * n_layers + [2]
        self.params = []
        for i in range(n_layers + 1):
            conv = torch.nn.Conv2d(dims[i], dims[i + 1], (k if i == 0 else 1, 1))        # actor.py:30-40 default init
            self.params += [conv.weight.detach().to(self.device).contiguous(), conv.bias.detach().to(self.device).contiguous()]
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.adam_step = 0
        self._push_weights()
        M = n_agents * n_episodes
        self.buffer_steps = buffer_steps
        self.z_buf = torch.zeros((buffer_steps, k, M, 6), dtype=torch.float32, device=self.device)
        self.label_buf = torch.zeros((buffer_steps, M, 2), dtype=torch.float32, device=self.device)
        self.filled, self.cursor = 0, 0
        self.beta = 1.0
        self._u_exp = torch.empty((M, 2), dtype=torch.float32, device=self.device)
        self._a_pol = torch.empty((M, 2), dtype=torch.float32, device=self.device)

    # -- weights ----------------------------------------------------------------------------
    def state_dict(self):
        return {f"conv_layers.{i // 2}.{'weight' if i % 2 == 0 else 'bias'}": p for i, p in enumerate(self.params)}

    def load_state_dict(self, sd):
        for i, p in enumerate(self.params):
            src = sd[f"conv_layers.{i // 2}.{'weight' if i % 2 == 0 else 'bias'}"]
            p.copy_(self.torch.as_tensor(np.asarray(src) if not hasattr(src, "to") else src).to(self.device).reshape(p.shape))
        self._push_weights()

    def _push_weights(self):
        self.engine.load_state_dict(self.state_dict())

    # -- one training episode (gnn_dagger.py:146-188) -----------------------------------------
    def run_episode(self, x0, steps, updates):
        """Roll B episodes for ``steps`` steps from x0 (B*N,4), storing every state; then ``updates`` gradient steps.
        Returns (mean per-episode return, summed policy loss)."""
        torch = self.torch
        eng = self.engine
        self.beta = max(self.beta * self.beta_coeff, 0.5)
        eng.reset(x0)
        ret = np.zeros(self.n_episodes)
        for _ in range(steps):
            eng.controller(centralized=self.centralized, max_accel=self.max_accel, out=self._u_exp)    # optimal_action
            eng.policy(out=self._a_pol)                                                                # select_action
            slot = self.cursor
            self.z_buf[slot].copy_(eng.get_aggregated(device=True))
            self.label_buf[slot].copy_(self._u_exp)
            self.cursor = (self.cursor + 1) % self.buffer_steps
            self.filled = min(self.filled + 1, self.buffer_steps)
            # np.random.binomial(1, beta) per episode (gnn_dagger.py:158): expert drives with probability beta
            expert = torch.rand(self.n_episodes, generator=self.gen, device=self.device) < self.beta
            mask = expert.repeat_interleave(self.n_agents).unsqueeze(1)
            action = torch.where(mask, self._u_exp, self._a_pol).contiguous()
            ret += eng.integrate(action, want_reward=True)
            eng.build_graph(advance=True)
        loss_sum = 0.0
        if self.filled * self.n_episodes > self.batch_size:
            for _ in range(updates):
                loss_sum += self.gradient_step()
            self._push_weights()                  # the rollout engine sees the new actor from the next episode on
        return float(ret.mean()), loss_sum

    def sample_batch(self):
        """``batch_size`` stored (step, episode) states: z (batch,K,N,6), labels (batch,2,N)."""
        torch = self.torch
        n_states = self.filled * self.n_episodes
        pick = torch.randint(0, n_states, (self.batch_size,), generator=self.gen, device=self.device)
        step, ep = pick // self.n_episodes, pick % self.n_episodes
        N = self.n_agents
        zb = self.z_buf[:self.filled].view(self.filled, self.k, self.n_episodes, N, 6)
        z = zb[step, :, ep]                                               # (batch,K,N,6)
        lb = self.label_buf[:self.filled].view(self.filled, self.n_episodes, N, 2)
        y = lb[step, ep].transpose(1, 2)                                  # (batch,2,N): gnn_dagger.py:173-175
        return z.contiguous(), y.contiguous()

    def gradient_step(self):
        z, y = self.sample_batch()
        self.adam_step += 1
        loss, _ = self.trainer.step(z, y, self.params, self.exp_avg, self.exp_avg_sq, step=self.adam_step, lr=self.lr)
        return float(loss.item())

    def close(self):
        self.engine.close()
        self.trainer.close()
