/*
 * bcmpc.h - C ABI of the B200-native random-shooting MPC planner.
 *
 * This is the drop-in boundary for the planning hot path of Bonnieccc/bc-mpc.  The
 * reference has no FFI of its own (the path sits behind duck-typed Python objects), so
 * each entry point cites the reference interface it replaces.  Plain pointers and sizes
 * only; no torch / numpy types.  Every function returns 0 on success or a negative
 * bcmpc_status; none throws across the ABI; bcmpc_last_error() returns the message.
 *
 * Threading: a handle is not re-entrant (one in-flight call per handle); use one handle per
 * device / rank.  Ownership: the caller owns every buffer it passes; set_* calls COPY their
 * inputs to the device (weights change after each dyn_model.fit, train_mpc_ppo.py:279, and
 * each PPO update, :355-364, so call them again after training steps).
 *
 * Tensor layouts follow the reference: weights are TF "kernel" layout [in, out] row-major
 * float32 with bias [out] (names NNDynamicsModel/dense{,_1,_2}/{kernel,bias}); action paths /
 * injected randomness are [H, K, A] row-major (controllers.py:53), per root state
 * ([B, H, K, A] when batched); states are float64 [*, D] (controllers.py:63).
 */
#ifndef BCMPC_H_
#define BCMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCMPC_ABI_VERSION 4   /* bumped whenever bcmpc_config, bcmpc_record or a signature changes */
#define BCMPC_MAX_ACT 8      /* action dims carried in a record (HalfCheetah: 6) */
#define BCMPC_MAX_OBS 32     /* observation dims supported (HalfCheetah: 20)      */
#define BCMPC_MAX_LAYERS 5   /* dense layers per MLP, output layer included        */

typedef struct bcmpc_handle bcmpc_handle;

typedef enum {
  BCMPC_OK = 0,
  BCMPC_ERR_INVALID = -1,     /* bad argument / configuration               */
  BCMPC_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail) */
  BCMPC_ERR_STATE = -3,       /* weights / normalisation not set yet          */
  BCMPC_ERR_UNSUPPORTED = -4  /* valid request the selected kernel cannot run */
} bcmpc_status;

/* hidden activation of the dynamics MLP (dynamics.py:66-67; all call sites use relu) */
enum { BCMPC_ACT_RELU = 0, BCMPC_ACT_TANH = 1 };

/* precision mode (BASELINE.json north_star): FP32_VERIFY follows the reference precision map
 * of dynamics.py:106-119 (float64 normalise / state carry / cost, float32 MLP on FFMA);
 * BF16_FAST is the tcgen05 path (bf16 operands, fp32 accumulate, fp32 state and cost). */
enum { BCMPC_PREC_FP32_VERIFY = 0, BCMPC_PREC_BF16_FAST = 1 };

/* action proposal mode:
 *  P0 UNIFORM           controllers.py:43-55,70     a = inj,               inj ~ U[low,high)
 *  P1 POLICY_GAUSSIAN   controllers.py:203          a = mean(s)+exp(logstd)*inj,  inj ~ N(0,1)
 *  P2 POLICY_MIX        controllers.py:205-208      a = (1-e)*mean(s)+e*inj,      inj ~ U[low,high)
 *  P3 NET_PLUS_UNIFORM  controllers.pyc src 120-163 a = net(s)+inj,               inj ~ U[-1,1)
 *  P4 POLICY_MEAN       controllers.py:429          a = mean(s)  (act(stochastic=False)), no randomness
 */
enum {
  BCMPC_PROP_UNIFORM = 0,
  BCMPC_PROP_POLICY_GAUSSIAN = 1,
  BCMPC_PROP_POLICY_MIX = 2,
  BCMPC_PROP_NET_PLUS_UNIFORM = 3,
  BCMPC_PROP_POLICY_MEAN = 4
};

/* Replaces the constructor arguments of MPCcontroller / MPCcontrollerPolicyNet
 * (controllers.py:28-41,162-181) and NNDynamicsModel (dynamics.py:8-52). */
typedef struct {
  int32_t obs_dim;      /* D  = env.observation_space.shape[0]                     */
  int32_t act_dim;      /* A  = env.action_space.shape[0]                          */
  int32_t n_hidden;     /* n_layers of NNDynamicsModel (hidden layers)              */
  int32_t hidden_size;  /* size                                                     */
  int32_t activation;   /* BCMPC_ACT_*                                              */
  int32_t num_paths;    /* K paths THIS handle rolls per root state                 */
  int32_t horizon;      /* H                                                        */
  int32_t max_states;   /* B_max concurrent root states per plan call (>=1)         */
  int32_t precision;    /* BCMPC_PREC_*                                             */
  int32_t proposal;     /* BCMPC_PROP_*                                             */
  int32_t device;       /* CUDA device ordinal                                      */
  int32_t first_proposal; /* 0: every step uses `proposal`.  1 + BCMPC_PROP_*: step 0 uses that mode and steps
                           * 1..H-1 use `proposal` - the first-stage / rollout split of
                           * MCTScontrollerPolicyNetReward.get_action (controllers.py:397-457).  Verification mode. */
  int64_t path_offset;  /* global index of this handle's first path (GPU sharding)  */
  int64_t total_paths;  /* global K over all shards (0 -> num_paths)                */
  double explore;       /* e of P2 (MPCcontrollerPolicyNet.explore)                 */
} bcmpc_config;

/* One winning candidate per root state: what controllers.py:82-85 extracts. */
typedef struct {
  double cost;                    /* costs[argmin]          (controllers.py:83)     */
  int64_t index;                  /* global argmin index j  (controllers.py:82)     */
  double action[BCMPC_MAX_ACT];   /* action_paths[0, j, :]  (controllers.py:84-85)  */
} bcmpc_record;

int bcmpc_abi_version(void);

/* message of the last failing call on this handle (or on this thread when h == NULL) */
const char* bcmpc_last_error(const bcmpc_handle* h);

int bcmpc_create(const bcmpc_config* cfg, bcmpc_handle** out);
int bcmpc_destroy(bcmpc_handle* h);

/* NNDynamicsModel weights (dynamics.py:54-71).  n_layers = n_hidden + 1; W[i] is [in_i,out_i]. */
int bcmpc_set_dynamics(bcmpc_handle* h, const float* const* W, const float* const* b, int32_t n_layers);

/* The same hand-off for a trainer whose parameters already live on this GPU (the PyTorch successor of
 * dyn_model.fit, dynamics.py:81-104, train_mpc_ppo.py:279): W[i] / b[i] are DEVICE pointers (float32, contiguous);
 * layout BCMPC_LAYOUT_IN_OUT = [in_i,out_i] like the TF kernels, BCMPC_LAYOUT_OUT_IN = [out_i,in_i] like
 * torch.nn.Linear.weight (transposed on the fly).  The copy and the bf16 re-pack of the tensor-core images are enqueued on `stream`
 * (NULL = the handle's stream) and the call returns without a host round trip; later plans on any stream are ordered
 * after it.  The caller must not have plans of this handle in flight on OTHER streams while it runs.  Resets
 * LayerNorm like bcmpc_set_dynamics. */
enum { BCMPC_LAYOUT_IN_OUT = 0, BCMPC_LAYOUT_OUT_IN = 1 };
int bcmpc_set_dynamics_device(bcmpc_handle* h, const float* const* dW, const float* const* db, int32_t n_layers,
                              int32_t layout, void* stream);

/* LayerNorm variant of NNDynamicsModel (dynamics.py:68-69, FLAGS.LAYER_NORM, the default of train_mpc_ppo.py:52):
 * tf.contrib.layers.layer_norm AFTER the activation of every hidden layer - moments over the last axis, centre and
 * scale.  gamma[l], beta[l] are [hidden_size] for hidden layer l (n_hidden of them); eps is the variance epsilon
 * (1e-12 in tf.contrib).  gamma == NULL switches LayerNorm off.  Call it after bcmpc_set_dynamics (new weights reset
 * it).  Both precision modes; with BCMPC_PREC_BF16_FAST: ReLU or tanh dynamics of two hidden layers (<= 510 units; also
 * together with a proposal network of <= 128 units, ReLU: <= 256) and the tanh learned-reward model, where a reward head needs
 * bcmpc_set_reward_layer_norm too (all three LayerNorms or none), and the cluster-split kernel (n_hidden > 2 or
 * hidden_size > 510; ReLU or tanh). */
int bcmpc_set_layer_norm(bcmpc_handle* h, const float* const* gamma, const float* const* beta, int32_t n_hidden, float eps);
/* ... with DEVICE pointers, enqueued on `stream` (see bcmpc_set_dynamics_device). */
int bcmpc_set_layer_norm_device(bcmpc_handle* h, const float* const* d_gamma, const float* const* d_beta, int32_t n_hidden,
                                float eps, void* stream);

/* Learned-reward planners: NNDynamicsRewardModel (dynamics.py:121-238) = the dynamics net (n_hidden = 2: shared layer,
 * delta-head hidden layer, delta output) plus a reward head that branches off the shared layer: W[0] [hidden,hidden]
 * with the dynamics activation, W[1] [hidden,1] linear.  With a reward head the planner maximises the sum of
 * gamma^i * (y_r * std_reward + mean_reward) (MPCcontrollerReward.get_action, controllers.py:121-158; gamma = 1 gives
 * MPCcontrollerPolicyNetReward, :319-363): bcmpc_record.cost is MINUS that sum, index / action its argmax (first index
 * on ties, like np.argmax).  W == NULL switches back to cheetah_cost_fn.  Call after bcmpc_set_dynamics; the head
 * persists across later bcmpc_set_dynamics calls.  BCMPC_PREC_BF16_FAST runs it on the tensor cores for the
 * reference's activation (tanh, dynamics.py:162) at every hidden size <= 510 (both heads form one 2 x hidden wide second
 * layer), with uniform or policy-guided proposals; ReLU heads fail with BCMPC_ERR_UNSUPPORTED there and run in
 * verification mode. */
int bcmpc_set_reward_head(bcmpc_handle* h, const float* const* W, const float* const* b, double mean_reward,
                          double std_reward, double gamma);

/* LayerNorm after the reward head's hidden layer (dynamics.py:175-176, FLAGS.LAYER_NORM): gamma, beta [hidden], or NULL
 * to switch it off.  The shared layer and the delta head take theirs from bcmpc_set_layer_norm.  Call after
 * bcmpc_set_reward_head (a new head resets it). */
int bcmpc_set_reward_layer_norm(bcmpc_handle* h, const float* gamma, const float* beta, float eps);

/* The reward head and its LayerNorm from DEVICE pointers (layout as in bcmpc_set_dynamics_device), enqueued on `stream`:
 * together with bcmpc_set_dynamics_device / bcmpc_set_layer_norm_device a PyTorch trainer refreshes the whole
 * NNDynamicsRewardModel after fit (dynamics.py:195-219) without a host round trip. */
int bcmpc_set_reward_head_device(bcmpc_handle* h, const float* const* dW, const float* const* db, int32_t layout,
                                 double mean_reward, double std_reward, double gamma, void* stream);
int bcmpc_set_reward_layer_norm_device(bcmpc_handle* h, const float* d_gamma, const float* d_beta, float eps, void* stream);

/* NNDynamicsRewardModel.predict (dynamics.py:221-238): next_states [n,D] and rewards [n].  HOST buffers. */
int bcmpc_predict_reward(bcmpc_handle* h, const double* states, const double* actions, int64_t n, double* next_states,
                         double* rewards);

/* The six members of the normalisation 10-tuple that predict uses (dynamics.py:41,109-113). */
int bcmpc_set_normalization(bcmpc_handle* h, const double* mean_obs, const double* std_obs,
                            const double* mean_act, const double* std_act,
                            const double* mean_delta, const double* std_delta);

/* env.action_space.low / high (controllers.py:53). */
int bcmpc_set_action_bounds(bcmpc_handle* h, const double* low, const double* high);

/* Proposal network: MlpPolicy 'pol' branch (ppo_bc_policy.py:60,70-72,79) or BCnetwork
 * (behavioral_cloning.py:21-26).  sizes has n_layers+1 entries (sizes[0] = D, last = A); hidden
 * layers are tanh, the last is linear.  logstd [A] may be NULL (P2/P3).  filt_mean/filt_std [D]
 * (RunningMeanStd mean/std as float32) may be NULL = no filter; clip <= 0 = no clipping.
 * BCMPC_PREC_BF16_FAST: two hidden layers of at most 256 units (MlpPolicy hid_size 128 / 256, BCnetwork 128 / 64); a network
 * with more than 128 units in a layer runs on kernel variants with a smaller weight ring, and switching between the two
 * classes re-packs the dynamics images; next to a dynamics net on the cluster-split kernel (n_hidden > 2 or hidden_size > 510,
 * which runs policy-guided shooting at 129..1024 hidden units) at most 128 units per layer.  Wider / deeper networks:
 * BCMPC_ERR_UNSUPPORTED (verification mode takes any). */
int bcmpc_set_policy(bcmpc_handle* h, int32_t n_layers, const int32_t* sizes, const float* const* W,
                     const float* const* b, const float* logstd, const float* filt_mean,
                     const float* filt_std, float clip);

/* The same hand-off after a PPO / BC update (ppo_bc_policy.py:135-151, train_mpc_ppo.py:355-364) when the trainer's
 * parameters live on this GPU: dW[i] / db[i] are DEVICE pointers (float32, contiguous, layout as in
 * bcmpc_set_dynamics_device); copy and re-pack are enqueued on `stream` (NULL = the handle's) and later plans on any
 * stream are ordered after them.  logstd, filt_mean, filt_std (a few floats the host-side RunningMeanStd / optimiser
 * own) stay HOST pointers. */
int bcmpc_set_policy_device(bcmpc_handle* h, int32_t n_layers, const int32_t* sizes, const float* const* dW,
                            const float* const* db, int32_t layout, const float* logstd, const float* filt_mean,
                            const float* filt_std, float clip, void* stream);

/* MPCcontroller.get_action / MPCcontrollerPolicyNet.get_action (controllers.py:57-88,189-237)
 * for B root states at once, HOST buffers.  states [B,D]; inj [B,H,K,A] or NULL (NULL = sample
 * on the device with Philox4x32-10 keyed by seed and the global path index).  out [B]. */
int bcmpc_plan(bcmpc_handle* h, const double* states, int32_t n_states, const double* inj,
               uint64_t seed, bcmpc_record* out);

/* Same, DEVICE buffers, asynchronous on `stream` (a cudaStream_t, NULL = the handle's own). */
int bcmpc_plan_device(bcmpc_handle* h, const double* d_states, int32_t n_states, const double* d_inj,
                      uint64_t seed, bcmpc_record* d_out, void* stream);

/* Same, with the Philox key read from DEVICE memory (*d_seed) when the kernels run instead of being a launch argument.
 * The call makes no host round trip and is capturable (cudaStreamBeginCapture / torch.cuda.graph; make one un-captured
 * call first: the first launch of a kernel sets its attributes): the vectorised env loop of utils.py:193-213 - policy /
 * env step / plan per time step - becomes one graph, and every replay draws new action sequences once the caller has
 * advanced *d_seed (e.g. a one-element tensor incremented inside the same graph).  Same results as bcmpc_plan_device
 * with seed = *d_seed. */
int bcmpc_plan_device_seeded(bcmpc_handle* h, const double* d_states, int32_t n_states, const double* d_inj,
                             const uint64_t* d_seed, bcmpc_record* d_out, void* stream);

/* Verification only: one root state, every path's states [H+1,K,D] (controllers.py:74), actions
 * [H,K,A] and trajectory costs [K] (controllers.py:80).  Any output may be NULL.  HOST buffers. */
int bcmpc_rollout_debug(bcmpc_handle* h, const double* state, const double* inj, uint64_t seed,
                        double* states_out, double* actions_out, double* costs_out);

/* NNDynamicsModel.predict (dynamics.py:106-119): n independent (state, action) rows. HOST buffers. */
int bcmpc_predict(bcmpc_handle* h, const double* states, const double* actions, int64_t n, double* next_states);

/* policy_net.act(ob, stochastic) (ppo_bc_policy.py:174-185) / BCnetwork.predict: noise [n,A] N(0,1)
 * draws for the stochastic sample, NULL = mode (mean).  out float32 [n,A].  HOST buffers. */
int bcmpc_policy_act(bcmpc_handle* h, const double* obs, const float* noise, int64_t n, float* out);

/* cheetah_cost_fn batched branch (cost_functions.py:10-30) on the device.  HOST buffers. */
int bcmpc_cheetah_cost(bcmpc_handle* h, const double* states, const double* next_states, int64_t n, double* out);

/* The injection tensor [B,H,K,A] the planner would draw on the device for (seed, proposal). */
int bcmpc_sample_injection(bcmpc_handle* h, uint64_t seed, int32_t n_states, double* out);

/* Cross-GPU min-loc: gathered [n_ranks, n_states] records -> [n_states] winners, lowest global
 * index on ties, NaN wins (np.argmin, controllers.py:82).  DEVICE buffers, asynchronous. */
int bcmpc_reduce_records(bcmpc_handle* h, const bcmpc_record* d_gathered, int32_t n_ranks,
                         int32_t n_states, bcmpc_record* d_out, void* stream);

/* The calling thread's current CUDA device: what the Python mirror's helper planners (NNDynamicsModel.predict,
 * MlpPolicy.act, cheetah_cost_fn) use when no device is given - never a hard-wired device 0.  Every other entry
 * point runs on its handle's device and RESTORES the caller's current device before it returns. */
int bcmpc_current_device(int32_t* device);

/* Cross-GPU min-loc through NVLink peer memory instead of ncclAllGather + bcmpc_reduce_records (the
 * one exchange step of a sharded plan).  peer_bufs[r] is rank r's exchange buffer as THIS process sees it (peer-mapped
 * device pointers, e.g. the buffer_ptrs of a torch symmetric-memory rendezvous; peer_bufs[rank] is local), each of
 * bcmpc_peer_exchange_bytes(world, max_states) bytes and zero-filled once before the first plan.  After this call every
 * bcmpc_plan_device of the handle is a COLLECTIVE over the `world` handles of the group: one rollout launch plus one
 * exchange kernel that stores this rank's per-state winners into every peer's buffer, flags them, waits for the peers'
 * and reduces (lowest global index on ties, np.argmin, controllers.py:82) - d_out holds the GLOBAL winners, identical
 * on every rank.  peer_bufs == NULL switches back to the local result.  world <= 16. */
int64_t bcmpc_peer_exchange_bytes(int32_t world, int32_t max_states);
int bcmpc_set_peer_exchange(bcmpc_handle* h, int32_t world, int32_t rank, void* const* peer_bufs);

/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t bcmpc_kernel_launches(const bcmpc_handle* h);

/* device time (ms, CUDA events on the launch stream) of the last bcmpc_plan* rollout kernel; the
 * events are recorded around the kernel only when timing was enabled. */
int bcmpc_enable_timing(bcmpc_handle* h, int32_t on);
int bcmpc_last_kernel_ms(bcmpc_handle* h, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* BCMPC_H_ */
