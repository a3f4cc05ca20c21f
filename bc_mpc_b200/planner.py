"""Planner: Python owner of one bcmpc handle (one GPU).  Thin: argument checking, numpy <-> pointers."""
import ctypes as C

import numpy as np

from . import _lib as L

_PROPOSALS = {"uniform": L.PROP_UNIFORM, "policy_gaussian": L.PROP_POLICY_GAUSSIAN,
              "policy_mix": L.PROP_POLICY_MIX, "bc_uniform": L.PROP_NET_PLUS_UNIFORM,
              "policy_mean": L.PROP_POLICY_MEAN}
_PRECISIONS = {"fp32": L.PREC_FP32_VERIFY, "fp32_verify": L.PREC_FP32_VERIFY,
               "bf16": L.PREC_BF16_FAST, "bf16_fast": L.PREC_BF16_FAST}
_ACTS = {"relu": L.ACT_RELU, "tanh": L.ACT_TANH}


_CUDA_STREAM_LEGACY = 1   # cudaStreamLegacy: the explicit handle of the legacy default stream (stream 0)


def _stream(stream_ptr):
    if stream_ptr is None:
        return None
    return C.c_void_p(int(stream_ptr) or _CUDA_STREAM_LEGACY)


def _handoff_stream(stream_ptr, tensor):
    """Stream of a device-pointer weight hand-off.  None = the stream the producer of `tensor` is (by convention) running on:
    torch's CURRENT stream of the tensor's device, so that the copy is ordered after the optimiser step that wrote the
    parameters without any synchronize; for non-torch tensors the handle's own stream (the caller synchronises)."""
    if stream_ptr is None and type(tensor).__module__.split(".")[0] == "torch":
        import torch
        stream_ptr = torch.cuda.current_stream(tensor.device).cuda_stream
    return _stream(stream_ptr)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


class Planner:
    """One random-shooting planner instance on one device.

    Mirrors what MPCcontroller / MPCcontrollerPolicyNet hold (controllers.py:28-41,162-181):
    K = num_paths, H = horizon, the dynamics net and the optional proposal net."""

    def __init__(self, obs_dim=20, act_dim=6, n_hidden=2, hidden_size=500, activation="relu", num_paths=1000,
                 horizon=15, max_states=1, precision="fp32", proposal="uniform", explore=1.0, device=0,
                 path_offset=0, total_paths=0, first_proposal=None):
        self._h = C.c_void_p()
        self._lib = L.lib()
        cfg = L.Config(obs_dim=obs_dim, act_dim=act_dim, n_hidden=n_hidden, hidden_size=hidden_size,
                       activation=_ACTS[activation] if isinstance(activation, str) else int(activation),
                       num_paths=num_paths, horizon=horizon, max_states=max_states,
                       precision=_PRECISIONS[precision] if isinstance(precision, str) else int(precision),
                       proposal=_PROPOSALS[proposal] if isinstance(proposal, str) else int(proposal),
                       device=device, first_proposal=0 if first_proposal is None else 1 + _PROPOSALS[first_proposal], path_offset=path_offset, total_paths=total_paths,
                       explore=float(explore))
        L.check(self._lib.bcmpc_create(C.byref(cfg), C.byref(self._h)))
        self.cfg = cfg
        self.D, self.A, self.K, self.H = obs_dim, act_dim, num_paths, horizon
        self.max_states = max_states
        self.precision = precision
        self.proposal = proposal

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bcmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        L.check(rc, self._h)

    # -- weights ----------------------------------------------------------------------------------
    def set_dynamics(self, weights):
        """weights: [(kernel[in,out], bias[out]), ...] hidden layers first (dynamics.py:54-71)."""
        Ws = [np.ascontiguousarray(w, np.float32) for w, _ in weights]
        bs = [np.ascontiguousarray(b, np.float32).reshape(-1) for _, b in weights]
        sizes = [self.D + self.A] + [self.cfg.hidden_size] * self.cfg.n_hidden + [self.D]
        if len(Ws) != len(sizes) - 1:
            raise ValueError(f"expected {len(sizes) - 1} layers, got {len(Ws)}")
        for i, (w, b) in enumerate(zip(Ws, bs)):
            if w.shape != (sizes[i], sizes[i + 1]) or b.shape != (sizes[i + 1],):
                raise ValueError(f"layer {i}: expected kernel {(sizes[i], sizes[i + 1])}, got {w.shape} / {b.shape}")
        self._ck(self._lib.bcmpc_set_dynamics(self._h, L.ptr_array(Ws), L.ptr_array(bs), len(Ws)))

    def set_dynamics_device(self, weights, layout="in_out", stream_ptr=None):
        """Weight hand-off without a host round trip (bcmpc_set_dynamics_device): weights = [(W, b), ...] of float32
        CUDA tensors on this planner's device (anything with data_ptr() / shape / is_contiguous(), e.g. torch
        parameters), W [in,out] (layout="in_out", TF kernels) or [out,in] (layout="out_in", torch.nn.Linear.weight).
        Enqueued on stream_ptr (None = torch's current stream of the tensors' device, i.e. ordered after the kernels that
        produced them; the handle's stream for non-torch tensors); the tensors must stay alive until that stream has run."""
        sizes = [self.D + self.A] + [self.cfg.hidden_size] * self.cfg.n_hidden + [self.D]
        if len(weights) != len(sizes) - 1:
            raise ValueError(f"expected {len(sizes) - 1} layers, got {len(weights)}")
        if layout not in ("in_out", "out_in"):
            raise ValueError("layout must be 'in_out' or 'out_in'")
        for i, (w, b) in enumerate(weights):
            want = (sizes[i], sizes[i + 1]) if layout == "in_out" else (sizes[i + 1], sizes[i])
            if tuple(w.shape) != want or tuple(b.shape) != (sizes[i + 1],):
                raise ValueError(f"layer {i}: expected kernel {want} / bias {(sizes[i + 1],)}, got {tuple(w.shape)} / {tuple(b.shape)}")
            for t in (w, b):
                if not t.is_contiguous() or "float32" not in str(t.dtype) or not t.is_cuda:
                    raise ValueError(f"layer {i}: tensors must be contiguous float32 CUDA tensors")
        Wp = (C.c_void_p * len(weights))(*[w.data_ptr() for w, _ in weights])
        bp = (C.c_void_p * len(weights))(*[b.data_ptr() for _, b in weights])
        self._ck(self._lib.bcmpc_set_dynamics_device(self._h, Wp, bp, len(weights), 0 if layout == "in_out" else 1,
                                                     _handoff_stream(stream_ptr, weights[0][0])))

    def set_layer_norm(self, params, eps=1e-12):
        """LayerNorm after every hidden activation (dynamics.py:68-69): params = [(gamma[size], beta[size]), ...] per
        hidden layer, or None to switch it off.  Call after set_dynamics (new weights reset it)."""
        if params is None:
            self._ck(self._lib.bcmpc_set_layer_norm(self._h, None, None, 0, C.c_float(eps)))
            return
        gs = [np.ascontiguousarray(g, np.float32).reshape(-1) for g, _ in params]
        bs = [np.ascontiguousarray(b, np.float32).reshape(-1) for _, b in params]
        for g, b in zip(gs, bs):
            if g.shape != (self.cfg.hidden_size,) or b.shape != (self.cfg.hidden_size,):
                raise ValueError(f"LayerNorm gamma/beta must be [{self.cfg.hidden_size}]")
        self._ck(self._lib.bcmpc_set_layer_norm(self._h, L.ptr_array(gs), L.ptr_array(bs), len(gs), C.c_float(eps)))

    @staticmethod
    def _check_cuda_f32(t, shape, what):
        if tuple(t.shape) != tuple(shape) or not t.is_contiguous() or "float32" not in str(t.dtype) or not t.is_cuda:
            raise ValueError(f"{what}: expected a contiguous float32 CUDA tensor of shape {tuple(shape)}, got {tuple(t.shape)}")

    def set_layer_norm_device(self, params, eps=1e-12, stream_ptr=None):
        """set_layer_norm with [(gamma, beta), ...] float32 CUDA tensors, enqueued on stream_ptr (bcmpc_set_layer_norm_device)."""
        S = self.cfg.hidden_size
        for g, b in params:
            self._check_cuda_f32(g, (S,), "LayerNorm gamma")
            self._check_cuda_f32(b, (S,), "LayerNorm beta")
        gp = (C.c_void_p * len(params))(*[g.data_ptr() for g, _ in params])
        bp = (C.c_void_p * len(params))(*[b.data_ptr() for _, b in params])
        self._ck(self._lib.bcmpc_set_layer_norm_device(self._h, gp, bp, len(params), C.c_float(eps), _handoff_stream(stream_ptr, params[0][0])))

    def set_reward_head_device(self, weights, mean_reward=0.0, std_reward=1.0, gamma=1.0, layer_norm=None, eps=1e-12, layout="in_out",
                               stream_ptr=None):
        """set_reward_head with float32 CUDA tensors ([(W[S,S], b[S]), (W[S,1], b[1])], or [out,in] kernels with
        layout="out_in"; layer_norm = (gamma, beta) tensors), enqueued on stream_ptr."""
        S = self.cfg.hidden_size
        shapes = [(S, S), (S, 1)] if layout == "in_out" else [(S, S), (1, S)]
        for (w, b), ws, bs in zip(weights, shapes, [(S,), (1,)]):
            self._check_cuda_f32(w, ws, "reward-head kernel")
            self._check_cuda_f32(b, bs, "reward-head bias")
        Wp = (C.c_void_p * 2)(*[w.data_ptr() for w, _ in weights])
        bp = (C.c_void_p * 2)(*[b.data_ptr() for _, b in weights])
        self._ck(self._lib.bcmpc_set_reward_head_device(self._h, Wp, bp, 0 if layout == "in_out" else 1, float(mean_reward),
                                                        float(std_reward), float(gamma), _handoff_stream(stream_ptr, weights[0][0])))
        if layer_norm is not None:
            self._check_cuda_f32(layer_norm[0], (S,), "reward-head LayerNorm gamma")
            self._check_cuda_f32(layer_norm[1], (S,), "reward-head LayerNorm beta")
            self._ck(self._lib.bcmpc_set_reward_layer_norm_device(self._h, C.c_void_p(layer_norm[0].data_ptr()),
                                                                  C.c_void_p(layer_norm[1].data_ptr()), C.c_float(eps), _handoff_stream(stream_ptr, layer_norm[0])))

    def set_reward_head(self, weights, mean_reward=0.0, std_reward=1.0, gamma=1.0, layer_norm=None, eps=1e-12):
        """Reward head of NNDynamicsRewardModel (dynamics.py:175-178): weights = [(kernel[S,S], bias[S]), (kernel[S,1], bias[1])],
        or None to go back to cheetah_cost_fn.  The planner then maximises sum_i gamma**i * reward_i (controllers.py:121-158).
        layer_norm = (gamma[S], beta[S]) of the head's hidden layer when FLAGS.LAYER_NORM (dynamics.py:176)."""
        if weights is None:
            self._ck(self._lib.bcmpc_set_reward_head(self._h, None, None, 0.0, 1.0, 1.0))
            return
        Ws = [np.ascontiguousarray(w, np.float32) for w, _ in weights]
        bs = [np.ascontiguousarray(b, np.float32).reshape(-1) for _, b in weights]
        S = self.cfg.hidden_size
        if len(Ws) != 2 or Ws[0].shape != (S, S) or Ws[1].shape != (S, 1) or bs[0].shape != (S,) or bs[1].shape != (1,):
            raise ValueError(f"reward head must be [({S},{S}), ({S},1)] kernels")
        self._ck(self._lib.bcmpc_set_reward_head(self._h, L.ptr_array(Ws), L.ptr_array(bs), float(mean_reward),
                                                 float(std_reward), float(gamma)))
        if layer_norm is not None:
            g = np.ascontiguousarray(layer_norm[0], np.float32).reshape(-1)
            b = np.ascontiguousarray(layer_norm[1], np.float32).reshape(-1)
            if g.shape != (S,) or b.shape != (S,):
                raise ValueError(f"reward-head LayerNorm gamma/beta must be [{S}]")
            self._ck(self._lib.bcmpc_set_reward_layer_norm(self._h, L.ptr(g), L.ptr(b), C.c_float(eps)))

    def predict_reward(self, states, actions):
        """NNDynamicsRewardModel.predict (dynamics.py:221-238): (next_states [n,D], rewards [n,1])."""
        states, actions = _f64(states), _f64(actions)
        n = states.shape[0]
        if states.shape != (n, self.D) or actions.shape != (n, self.A):
            raise ValueError(f"predict expects states [n,{self.D}] and actions [n,{self.A}]")
        out, rew = np.empty_like(states), np.empty(n)
        self._ck(self._lib.bcmpc_predict_reward(self._h, L.ptr(states), L.ptr(actions), n, L.ptr(out), L.ptr(rew)))
        return out, rew.reshape(n, 1)

    def set_normalization(self, mean_obs, std_obs, mean_action, std_action, mean_deltas, std_deltas):
        a = [_f64(mean_obs, (self.D,)), _f64(std_obs, (self.D,)), _f64(mean_action, (self.A,)),
             _f64(std_action, (self.A,)), _f64(mean_deltas, (self.D,)), _f64(std_deltas, (self.D,))]
        self._ck(self._lib.bcmpc_set_normalization(self._h, *[L.ptr(x) for x in a]))

    def set_action_bounds(self, low, high):
        low = _f64(np.broadcast_to(low, (self.A,)))
        high = _f64(np.broadcast_to(high, (self.A,)))
        self._ck(self._lib.bcmpc_set_action_bounds(self._h, L.ptr(low), L.ptr(high)))

    def set_policy(self, weights, logstd=None, filt_mean=None, filt_std=None, clip=0.0):
        Ws = [np.ascontiguousarray(w, np.float32) for w, _ in weights]
        bs = [np.ascontiguousarray(b, np.float32).reshape(-1) for _, b in weights]
        sizes = np.array([Ws[0].shape[0]] + [w.shape[1] for w in Ws], np.int32)
        for i, (w, b) in enumerate(zip(Ws, bs)):
            if w.shape != (sizes[i], sizes[i + 1]) or b.shape != (sizes[i + 1],):
                raise ValueError(f"policy layer {i}: inconsistent shapes {w.shape} / {b.shape}")
        ls = None if logstd is None else np.ascontiguousarray(logstd, np.float32).reshape(-1)
        fm = None if filt_mean is None else np.ascontiguousarray(filt_mean, np.float32).reshape(-1)
        fs = None if filt_std is None else np.ascontiguousarray(filt_std, np.float32).reshape(-1)
        self._ck(self._lib.bcmpc_set_policy(self._h, len(Ws), L.ptr(sizes), L.ptr_array(Ws), L.ptr_array(bs),
                                            L.ptr(ls), L.ptr(fm), L.ptr(fs), float(clip or 0.0)))

    def set_policy_device(self, weights, logstd=None, filt_mean=None, filt_std=None, clip=0.0, layout="in_out", stream_ptr=None):
        """Proposal-network hand-off without a host round trip (bcmpc_set_policy_device): weights = [(W, b), ...] float32
        CUDA tensors, W [in,out] (layout="in_out") or [out,in] (layout="out_in", torch.nn.Linear.weight); logstd and the
        observation-filter statistics are small host arrays.  Enqueued on stream_ptr (None = torch's current stream)."""
        if layout not in ("in_out", "out_in"):
            raise ValueError("layout must be 'in_out' or 'out_in'")
        dims = [(tuple(w.shape) if layout == "in_out" else tuple(w.shape)[::-1]) for w, _ in weights]
        sizes = np.array([dims[0][0]] + [d[1] for d in dims], np.int32)
        for i, (w, b) in enumerate(weights):
            if dims[i] != (sizes[i], sizes[i + 1]) or tuple(b.shape) != (sizes[i + 1],):
                raise ValueError(f"policy layer {i}: inconsistent shapes {tuple(w.shape)} / {tuple(b.shape)}")
            for t in (w, b):
                if not t.is_contiguous() or "float32" not in str(t.dtype) or not t.is_cuda:
                    raise ValueError(f"policy layer {i}: tensors must be contiguous float32 CUDA tensors")
        ls = None if logstd is None else np.ascontiguousarray(logstd, np.float32).reshape(-1)
        fm = None if filt_mean is None else np.ascontiguousarray(filt_mean, np.float32).reshape(-1)
        fs = None if filt_std is None else np.ascontiguousarray(filt_std, np.float32).reshape(-1)
        Wp = (C.c_void_p * len(weights))(*[w.data_ptr() for w, _ in weights])
        bp = (C.c_void_p * len(weights))(*[b.data_ptr() for _, b in weights])
        self._ck(self._lib.bcmpc_set_policy_device(self._h, len(weights), L.ptr(sizes), Wp, bp, 0 if layout == "in_out" else 1,
                                                   L.ptr(ls), L.ptr(fm), L.ptr(fs), float(clip or 0.0), _handoff_stream(stream_ptr, weights[0][0])))

    # -- the hot path -----------------------------------------------------------------------------
    def plan(self, states, inj=None, seed=0):
        """get_action for B root states (host buffers).  Returns a structured array [B] of
        (cost, index, action[8]).  inj: [B,H,K,A] (or [H,K,A] for B = 1) injected randomness."""
        states = _f64(states)
        if states.ndim == 1:
            states = states[None]
        B = states.shape[0]
        if states.shape[1] != self.D:
            raise ValueError(f"states must be [B,{self.D}]")
        if inj is not None:
            inj = _f64(inj).reshape(B, self.H, self.K, self.A)
        out = np.zeros(B, L.RECORD_DTYPE)
        self._ck(self._lib.bcmpc_plan(self._h, L.ptr(states), B, L.ptr(inj), C.c_uint64(seed), L.ptr(out)))
        return out

    def plan_device(self, d_states_ptr, n_states, d_out_ptr, seed=0, d_inj_ptr=None, stream_ptr=None, d_seed_ptr=None):
        """Asynchronous launch on device buffers (raw pointers, e.g. torch.Tensor.data_ptr()).
        stream_ptr: None = the handle's own stream; an integer cudaStream_t otherwise (0, torch's default
        stream, is passed as cudaStreamLegacy so that it is not mistaken for "none").
        d_seed_ptr: device address of a uint64 / int64 scalar that holds the Philox key (bcmpc_plan_device_seeded) - the
        launch can then be captured into a CUDA graph and replayed with new draws; `seed` is ignored."""
        if d_seed_ptr is not None:
            self._ck(self._lib.bcmpc_plan_device_seeded(self._h, C.c_void_p(d_states_ptr), n_states,
                                                        C.c_void_p(d_inj_ptr) if d_inj_ptr else None, C.c_void_p(d_seed_ptr),
                                                        C.c_void_p(d_out_ptr), _stream(stream_ptr)))
            return
        self._ck(self._lib.bcmpc_plan_device(self._h, C.c_void_p(d_states_ptr), n_states,
                                             C.c_void_p(d_inj_ptr) if d_inj_ptr else None, C.c_uint64(seed),
                                             C.c_void_p(d_out_ptr), _stream(stream_ptr)))

    def rollout_debug(self, state, inj=None, seed=0, want_states=True, want_actions=True):
        state = _f64(state, (self.D,))
        if inj is not None:
            inj = _f64(inj, (self.H, self.K, self.A))
        so = np.empty((self.H + 1, self.K, self.D)) if want_states else None
        ao = np.empty((self.H, self.K, self.A)) if want_actions else None
        co = np.empty(self.K)
        self._ck(self._lib.bcmpc_rollout_debug(self._h, L.ptr(state), L.ptr(inj), C.c_uint64(seed), L.ptr(so),
                                               L.ptr(ao), L.ptr(co)))
        return so, ao, co

    def predict(self, states, actions):
        states = _f64(states)
        actions = _f64(actions)
        n = states.shape[0]
        if states.shape != (n, self.D) or actions.shape != (n, self.A):
            raise ValueError(f"predict expects states [n,{self.D}] and actions [n,{self.A}]")
        out = np.empty_like(states)
        self._ck(self._lib.bcmpc_predict(self._h, L.ptr(states), L.ptr(actions), n, L.ptr(out)))
        return out

    def policy_act(self, obs, noise=None):
        obs = _f64(obs)
        n = obs.shape[0]
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32).reshape(n, self.A)
        out = np.empty((n, self.A), np.float32)
        self._ck(self._lib.bcmpc_policy_act(self._h, L.ptr(obs), L.ptr(nz), n, L.ptr(out)))
        return out

    def cheetah_cost(self, states, next_states):
        states, next_states = _f64(states), _f64(next_states)
        out = np.empty(states.shape[0])
        self._ck(self._lib.bcmpc_cheetah_cost(self._h, L.ptr(states), L.ptr(next_states), states.shape[0], L.ptr(out)))
        return out

    def sample_injection(self, seed, n_states=1):
        out = np.empty((n_states, self.H, self.K, self.A))
        self._ck(self._lib.bcmpc_sample_injection(self._h, C.c_uint64(seed), n_states, L.ptr(out)))
        return out

    def peer_exchange_bytes(self, world):
        return int(self._lib.bcmpc_peer_exchange_bytes(world, self.max_states))

    def set_peer_exchange(self, world, rank, peer_ptrs):
        """Makes every later plan_device a collective over `world` handles: winners over ALL ranks' paths, exchanged through
        NVLink peer memory (bcmpc_set_peer_exchange).  peer_ptrs[r] = rank r's exchange buffer as this process sees it
        (zero-filled, peer_exchange_bytes(world) each); None switches it off."""
        if peer_ptrs is None:
            self._ck(self._lib.bcmpc_set_peer_exchange(self._h, 0, 0, None))
            return
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._ck(self._lib.bcmpc_set_peer_exchange(self._h, world, rank, arr))

    def reduce_records_device(self, d_gathered_ptr, n_ranks, n_states, d_out_ptr, stream_ptr=None):
        self._ck(self._lib.bcmpc_reduce_records(self._h, C.c_void_p(d_gathered_ptr), n_ranks, n_states,
                                                C.c_void_p(d_out_ptr), _stream(stream_ptr)))

    # -- bookkeeping ------------------------------------------------------------------------------
    @property
    def kernel_launches(self):
        return int(self._lib.bcmpc_kernel_launches(self._h))

    def enable_timing(self, on=True):
        self._ck(self._lib.bcmpc_enable_timing(self._h, int(on)))

    def last_kernel_ms(self):
        ms = C.c_float()
        self._ck(self._lib.bcmpc_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)
