/*
 * fgnn.h -- C ABI of the B200-native flocking-GNN rollout engine (libfgnn.so).
 *
 * The reference (katetolstaya/multiagent_gnn_policies) has no FFI: its boundary for the
 * hot path is a Python call surface.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference tree).  A Python host binds these
 * with ctypes (multiagent_gnn_policies_b200/engine.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; fgnn_last_error() gives text.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - data pointers marked [h/d] may be HOST or DEVICE pointers (UVA, cudaMemcpyDefault);
 *     copies are stream-ordered (truly asynchronous only for pinned host memory).
 *   - agents are indexed a = episode * n_agents + i, in the caller's order, everywhere.
 *   - a handle is not re-entrant; distinct handles are independent.  No host threads.
 */
#ifndef FGNN_H
#define FGNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fgnn_handle fgnn_handle;

typedef struct fgnn_config {
    int32_t n_agents;        /* N, agents per episode            (cfg key n_agents)           */
    int32_t n_episodes;      /* B, independent episodes (block-diagonal graph), >= 1          */
    int32_t k;               /* filter taps K, 1..4              (cfg key k)                  */
    int32_t n_states;        /* F, must be 6                     (cfg key n_states)           */
    int32_t n_actions;       /* A, must be 2                     (cfg key n_actions)          */
    int32_t hidden;          /* H, 1..128                        (cfg key hidden_size)        */
    int32_t n_layers;        /* hidden layers L, 1..4            (cfg key n_layers, default 2)*/
    int32_t mean_pooling;    /* 1: state_network = adj / max(deg,1)  (gym_flock default)      */
    int32_t half_accel_term; /* 1: p += v dt + a dt^2/2 ; 0: p += v dt                        */
    int32_t device;          /* CUDA device ordinal                                           */
    int32_t grid_dim;        /* cells per side of the wrapped cell grid, 0 = auto             */
    int32_t edge_capacity;   /* directed-edge capacity per agent (mean), 0 = auto (48)        */
    int32_t readout_mode;    /* 0 = auto, 1 = FFMA (CUDA cores), 2 = tensor cores (tcgen05, 3xTF32) */
    int32_t grid_dim_y;      /* cells along y, 0 = same as grid_dim                           */
    int32_t shard_lo;        /* multi-GPU: first agent this rank owns ...                     */
    int32_t shard_count;     /* ... and how many (0 = all: single-GPU)                        */
    int32_t ghost_capacity;  /* multi-GPU: max agents received from other ranks per step      */
    int32_t flags;           /* FGNN_FLAG_* bits, 0 = defaults                                */
    double  comm_radius;     /* R                                (cfg key comm_radius)        */
    double  dt;              /*                                  (cfg key dt)                 */
    double  action_scalar;   /* gym_flock gain, 10.0                                          */
} fgnn_config;

/* fgnn_config.flags */
#define FGNN_FLAG_CSR_TAIL_ONLY 1   /* keep CSR rows only for agents with more than 8 neighbours (the first 8 of every
                                       row always sit in the per-agent ELL head): fgnn_get_csr is then unavailable,
                                       every other entry point works unchanged.  Saves 4*deg bytes per agent-step. */

typedef struct fgnn_stats {
    int64_t step;            /* index t of the current graph (0 right after reset)            */
    int64_t n_edges;         /* directed edges of the current graph                           */
    int32_t overflow;        /* 1: an edge / halo capacity was exceeded at some point (results void); 2: a p2p halo wait timed out */
    int32_t grid_dim;
    int64_t n_cells;
    int64_t edge_capacity;   /* total directed-edge capacity                                   */
    int64_t n_ghosts;        /* sharded handle: agents received from other ranks in the last exchange */
} fgnn_stats;

const char* fgnn_last_error(void);
int fgnn_version(void);

/* Lifetime.  Replaces: gym.make + env.env.params_from_cfg (train.py:17-21) and
 * DAGGER.__init__/Actor.__init__ (learner/gnn_dagger.py:20-53, learner/actor.py:9-42). */
int fgnn_create(const fgnn_config* cfg, fgnn_handle** out);
int fgnn_destroy(fgnn_handle* h);

/* Actor weights, conv layout of the reference checkpoint (learner/actor.py:30-40):
 * layer 0: W (H, F, K) [= conv weight (out,in,step,1)], layers 1..L-1: W (H, H), layer L: W (A, H);
 * b (out,).  fp32, [h/d].  Replaces Actor.load_state_dict (learner/gnn_dagger.py:114-123). */
int fgnn_set_weights(fgnn_handle* h, int32_t layer, const float* W, const float* b, void* stream);

/* env.reset() (learner/gnn_dagger.py:150): install x (B*N,4) f64 [px,py,vx,vy] [h/d], clear the
 * K-deep history, t = 0, build graph + features of step 0. */
int fgnn_reset(fgnn_handle* h, const double* x_bn4, void* stream);

/* Teacher forcing for parity tests: overwrite the 4-d state, keep history, do not rebuild. */
int fgnn_set_state(fgnn_handle* h, const double* x_bn4, void* stream);

/* gym_flock compute_helpers (called by env.step/reset): radius adjacency (CSR), degrees,
 * 6-d relative features of the CURRENT state into history slot t.  advance != 0 first moves
 * t -> t+1 (what env.step does after integrating). */
int fgnn_build_graph(fgnn_handle* h, int32_t advance, void* stream);

/* First half of env.step(u) (learner/gnn_dagger.py:163): double-integrator update from
 * u (B*N,2) fp32 [h/d]; reward_b (B,) f64 [h/d] or NULL receives instant_cost per episode. */
int fgnn_integrate(fgnn_handle* h, const float* u_bn2, double* reward_b, void* stream);

/* env.step(u) = fgnn_integrate + fgnn_build_graph(advance=1). */
int fgnn_env_step(fgnn_handle* h, const float* u_bn2, double* reward_b, void* stream);

/* The same two calls for a float64 action: when beta-mixing picks the expert, the reference steps the env with the
 * controller's float64 array itself (learner/gnn_dagger.py:156-163 -- no fp32 cast on that branch). */
int fgnn_integrate_f64(fgnn_handle* h, const double* u_bn2, double* reward_b, void* stream);
int fgnn_env_step_f64(fgnn_handle* h, const double* u_bn2, double* reward_b, void* stream);

/* DAGGER.select_action(state) (learner/gnn_dagger.py:55-72) = MultiAgentStateWithDelay history
 * (learner/state_with_delay.py:44-53) + Actor.forward (learner/actor.py:45-86) on the sparse
 * history the engine keeps: K-hop aggregation + readout.  action_bn2 (B*N,2) fp32 [h/d]. */
int fgnn_policy(fgnn_handle* h, float* action_bn2, void* stream);

/* env.env.controller(centralized) (learner/gnn_dagger.py:156, learner/gnn_baseline.py:16): expert
 * potential-based action for the CURRENT state/graph, (B*N,2) fp32 [h/d] in action units (already
 * clipped to +-max_accel and divided by the gain).  centralized != 0: velocity term over all agents
 * of the episode, potential term inside its own cut-off (r^2 <= comm_radius). */
int fgnn_controller(fgnn_handle* h, int32_t centralized, double max_accel, float* u_bn2, void* stream);
/* ... and in float64, as gym_flock returns it (no rounding to fp32 between the controller and env.step). */
int fgnn_controller_f64(fgnn_handle* h, int32_t centralized, double max_accel, double* u_bn2, void* stream);

/* One closed-loop rollout step (learner/gnn_dagger.py:196-201, test_model.py:38-45):
 * select_action -> env.step(action).  action_bn2 / reward_b may be NULL. */
int fgnn_step(fgnn_handle* h, float* action_bn2, double* reward_b, void* stream);

/* T closed-loop steps, CUDA-graph replayed; reward_bt (T,B) f64 [h/d] or NULL. */
int fgnn_rollout(fgnn_handle* h, int32_t T, double* reward_bt, void* stream);

/* Actor.forward(delay_state, delay_gso) on dense tensors (learner/actor.py:45-86):
 * delay_state (B2,K,F,N2), delay_gso (B2,K,N2,N2), out (B2,1,A,N2); fp32 DEVICE pointers. */
int fgnn_actor_forward_dense(fgnn_handle* h, int32_t batch, int32_t n_agents, const float* delay_state,
                             const float* delay_gso, float* out, void* stream);

/* Actor.forward(delay_state, delay_gso) on dense tensors for ANY aggregation index and layer widths the reference's
 * constructor accepts (learner/actor.py:9-43: kernel (K,1) at layer ind_agg, (1,1) elsewhere; forward: actor.py:45-86;
 * ind_agg > 0 is what learner/gnn_ddpg.py:126 builds).  No engine handle: the caller passes the layer parameters.
 *   widths      n_layers + 1 ints: n_s, hidden widths ..., n_a
 *   W[l], b[l]  conv_layers[l].weight (widths[l+1], widths[l], step_l, 1) contiguous with step_l = K at l == ind_agg,
 *               1 elsewhere, and conv_layers[l].bias (widths[l+1]); fp32 DEVICE pointers (the array of pointers is host)
 *   delay_state (B,K,n_s,N), delay_gso (B,K,N,N), out (B,1,n_a,N); fp32 DEVICE pointers
 *   workspace   DEVICE scratch of fgnn_actor_general_workspace(...) bytes (that call returns -1 on bad shapes)
 * ind_agg outside 0..n_layers-1 is accepted only for K = 1 (otherwise K rows remain and the reference's final view
 * fails too).  Small-N compatibility surface (dense operators), not the rollout path. */
int64_t fgnn_actor_general_workspace(int32_t batch, int32_t n_agents, int32_t k, int32_t n_layers, const int32_t* widths);
int fgnn_actor_forward_general(int32_t device, int32_t batch, int32_t n_agents, int32_t k, int32_t n_layers,
                               const int32_t* widths, int32_t ind_agg, const float* const* W, const float* const* b,
                               const float* delay_state, const float* delay_gso, float* out, float* workspace,
                               void* stream);

/* Read-back (tests, small-N compatibility with the dense reference API). */
int fgnn_get_state(fgnn_handle* h, double* x_bn4, void* stream);
int fgnn_get_features(fgnn_handle* h, int32_t age, float* values_bn6, void* stream);       /* x_{t-age} */
int fgnn_get_degrees(fgnn_handle* h, int32_t age, int32_t* deg_bn, void* stream);
int fgnn_get_aggregated(fgnn_handle* h, float* z_k_bn_6, void* stream);   /* (K, B*N, 6) of the last policy call */
int fgnn_get_action(fgnn_handle* h, float* action_bn2, void* stream);
/* Row-normalised state_network of graph t-age as dense (B, N, N) fp32 (state_with_delay.py:35). */
int fgnn_export_network_dense(fgnn_handle* h, int32_t age, float* network_bnn, void* stream);
/* CSR of graph t-age: device pointers valid until the next build; rows are (row_start[a], deg[a]). */
int fgnn_get_csr(fgnn_handle* h, int32_t age, const uint32_t** row_start, const int32_t** deg,
                 const int32_t** cols, const float** src_scale);
int fgnn_get_stats(fgnn_handle* h, fgnn_stats* out, void* stream);   /* synchronises the stream */

/* ---- multi-GPU (n_episodes == 1; every rank holds full-size arrays, indices are global) ----
 * Rank r starts owning agents [shard_lo, shard_lo + shard_count) and is the only one that integrates what it
 * owns.  One closed-loop step on a rank =
 *   fgnn_shard_local_step (K-hop aggregation over the agents present on the rank; readout + integrator for
 *   the OWNED agents) -> fgnn_shard_pack (records of owned agents that lie inside another rank's window; hand-over
 *   decisions) -> all-gather of the buffers by the host (NCCL) -> fgnn_shard_unpack (ghosts, new owned agents)
 *   -> fgnn_build_graph(advance = 1);   or the two CUDA-graph halves fgnn_shard_step_begin / _end.
 * Buffers are DEVICE arrays of doubles: per rank (cap + 1) records of 6 doubles; record 0 is the header
 * [count, x_lo, x_hi, 0, 0, 0] (x-interval of what the rank owns), records 1..count are
 * [agent id, px, py, vx, vy, new owner or -1].  `windows` is a device array with every rank's [x_lo, x_hi] at
 * windows[q * window_stride + {0, 1}] (the headers of the previous gathered buffer can be passed directly).
 * Territories: rank q's strip is bounds[q] <= x - shift < bounds[q+1] (bounds: world + 1 HOST doubles, bounds[0]
 * = -inf, bounds[world] = +inf), shift grows by dshift per step (frame moving with the flock).  An owner hands an
 * agent over to the rank whose strip it entered by more than `margin` (no data moves: the receiver already holds
 * it as a ghost with valid history), allowed from step index `handover_after` on.  `depth` = halo depth of the
 * windows (strip +- depth united with the owned x-interval +- depth).
 * On a sharded handle fgnn_policy / fgnn_integrate use arrays of (shard_count + ghost_capacity) rows in
 * OWNED-LIST order (fgnn_shard_owned returns the list); fgnn_step / fgnn_rollout / fgnn_env_step are refused. */
int fgnn_shard_configure(fgnn_handle* h, const double* bounds, int32_t world, int32_t rank, double depth,
                         double margin, double dshift, int32_t handover_after);
int fgnn_shard_local_step(fgnn_handle* h, void* stream);
int fgnn_shard_pack(fgnn_handle* h, const double* windows, int64_t window_stride, double* send_buf, int32_t cap,
                    int32_t advance, void* stream);
int fgnn_shard_unpack(fgnn_handle* h, const double* recv_buf, int32_t cap, void* stream);
int fgnn_shard_step_begin(fgnn_handle* h, const double* windows, int64_t window_stride, double* send_buf, int32_t cap,
                          void* stream);
int fgnn_shard_step_end(fgnn_handle* h, const double* recv_buf, int32_t cap, void* stream);
int fgnn_shard_owned(fgnn_handle* h, int32_t* ids, int32_t* count, void* stream);   /* synchronises */
/* The whole sharded step as ONE CUDA graph with the halo all-gather inside it (ncclAllGather on the engine's own
 * communicator, enqueued on the caller's stream between the two halves above): no second stream, no host round
 * trip between the halves.  fgnn_comm_unique_id: rank 0 obtains the 128-byte id and distributes it by any means
 * (the Python host broadcasts it through torch.distributed); fgnn_comm_init is collective over the ranks.
 * recv_buf: world * (cap + 1) records; its headers, as the previous step left them, are this step's windows. */
int fgnn_comm_unique_id(void* id_128);
int fgnn_comm_init(fgnn_handle* h, const void* id_128, int32_t rank, int32_t world);
int fgnn_shard_step(fgnn_handle* h, double* send_buf, double* recv_buf, int32_t cap, void* stream);

/* Halo over peer-to-peer stores instead of a collective (one node, NVLink / NVSwitch).  Every rank owns an inbox
 * [2 halves][world senders][cap + 1 records] (+ one flag word per half and sender).  In step t the closed final kernel of
 * rank s stores the records rank r needs straight into r's inbox slot [t & 1][s] through a CUDA-IPC mapping; a small kernel
 * then publishes the header [count, x_lo, x_hi] and, fenced, the flag t + 1; r's graph waits for the flags of its peers and
 * unpacks.  Only ranks whose windows overlap exchange anything (neighbouring strips), the payload does not grow with the
 * world size, and a step is ONE CUDA graph with no host round trip.
 *   fgnn_p2p_alloc    after fgnn_shard_configure: allocate the inbox; ipc_handle_64 (64 bytes, may be NULL) receives its
 *                     cudaIpcMemHandle_t, *local_ptr (may be NULL) the device pointer
 *   fgnn_p2p_connect  handles = world x 64 bytes in rank order (ranks in other processes), or direct_ptrs = world device
 *                     pointers (ranks in this process); the own entry is ignored
 *   fgnn_p2p_seed     after the reset-time exchange (fgnn_shard_pack / all-gather / fgnn_shard_unpack): the gathered buffer
 *                     becomes both halves of the inbox (its headers are the first step's windows)
 *   fgnn_shard_step_p2p  one closed-loop step.  A peer that never signals makes the wait give up after ~30 s and raises
 *                     fgnn_stats.overflow = 2 instead of hanging the device.
 *   fgnn_shard_exchange_p2p  the halo exchange alone (pack into the peers' inboxes, flags, wait, unpack) for a step whose
 *                     integrator ran on its own (fgnn_integrate with a host action -- the reference-facing loop,
 *                     learner/gnn_dagger.py:196-201); follow with fgnn_build_graph(advance). */
int fgnn_p2p_alloc(fgnn_handle* h, int32_t world, int32_t rank, int32_t cap, void* ipc_handle_64, void** local_ptr);
int fgnn_p2p_connect(fgnn_handle* h, const void* handles, void* const* direct_ptrs);
int fgnn_p2p_seed(fgnn_handle* h, const double* gathered, void* stream);
int fgnn_shard_step_p2p(fgnn_handle* h, void* stream);
int fgnn_shard_exchange_p2p(fgnn_handle* h, int32_t advance, void* stream);

/* ---- environment variants named by the reference's cfgs (gym_flock, un-vendored) ----
 * FlockingLeader-v0 (cfg/dagger_leader.cfg:24): mask_bn (B*N,) bytes [h/d], 0 = leader -- the integrator ignores
 * its action (u * mask), in every integrator path including the fused closed-loop kernel.  NULL removes the mask.
 * FlockingStochastic-v0 (cfg/dagger_stoch.cfg:24, no `dt` key): the time step is drawn by the env every step;
 * fgnn_set_dt installs it for the following integrations (cached CUDA graphs are re-captured). */
int fgnn_set_agent_mask(fgnn_handle* h, const uint8_t* mask_bn, void* stream);
int fgnn_set_dt(fgnn_handle* h, double dt);

/* ---- imitation-learning update ----
 * DAGGER.gradient_step (learner/gnn_dagger.py:76-96) / Cloning.gradient_step: Actor.forward on a batch of stored
 * states, F.mse_loss against the expert actions, backward, Adam step (learner/gnn_dagger.py:49, torch defaults).
 * With ind_agg = 0 (gnn_dagger.py:43) only the readout is trainable and it sits behind the aggregation, so the
 * batch is given as the aggregated features of each stored state, z = delay_state @ delay_gso (actor.py:70):
 *   z_bkn6      (batch, K, N, 6)  fp32 DEVICE  (what fgnn_get_aggregated returns per state)
 *   target_b2n  (batch, 2, N)     fp32 DEVICE  (torch.cat(batch.action), gnn_dagger.py:86)
 *   params / exp_avg / exp_avg_sq: HOST arrays of 2(L+1) DEVICE pointers in the order W_0, b_0, ..., W_L, b_L, each
 *     tensor in the reference's conv layout (learner/actor.py:30-40) -- the torch parameters and the torch.optim.Adam
 *     state tensors themselves, updated in place.  step = 1-based count of this update (Adam bias correction).
 *   apply = 0: loss / gradients only, nothing is modified (exp_avg* may be NULL).
 *   loss_out (1,) fp32 [h/d] or NULL; grads_out (fgnn_trainer_param_count,) fp32 [h/d] or NULL: the gradients packed
 *     in the same order as `params`. */
typedef struct fgnn_trainer fgnn_trainer;
int fgnn_trainer_create(int32_t k, int32_t hidden, int32_t n_layers, int32_t device, fgnn_trainer** out);
int fgnn_trainer_destroy(fgnn_trainer* tr);
int32_t fgnn_trainer_param_count(fgnn_trainer* tr);
int64_t fgnn_trainer_launch_count(fgnn_trainer* tr);
int fgnn_trainer_step(fgnn_trainer* tr, int32_t batch, int32_t n_agents, const float* z_bkn6, const float* target_b2n,
                      float* const* params, float* const* exp_avg, float* const* exp_avg_sq, int64_t step, double lr,
                      double beta1, double beta2, double eps, int32_t apply, float* loss_out, float* grads_out,
                      void* stream);

/* One closed-loop step (same work as fgnn_step) with a CUDA event after every kernel: ms_out[i] is the
 * device time of kernel i, names_out (16 bytes each, may be NULL) its name.  For bench.py's roofline. */
int fgnn_profile_step(fgnn_handle* h, int32_t max_kernels, float* ms_out, char* names_out, int32_t* n_out,
                      void* stream);

/* Stream-ordered copy between any two host/device buffers (cudaMemcpyDefault) + stream sync; lets a
 * ctypes host read the device arrays fgnn_get_csr points at without binding the CUDA runtime. */
int fgnn_memcpy_sync(void* dst, const void* src, uint64_t bytes, void* stream);

/* Number of kernel launches issued by this handle so far (bench `gpu_launches`). */
int64_t fgnn_launch_count(fgnn_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* FGNN_H */
