"""Drop-in for NNDynamicsModel's inference side (dynamics.py:8-52,106-119).

The reference builds a TF1 graph and trains it with `fit`; training is out of scope for this path
(SURVEY.md section 8a, row a5), so the model here is a weight container + GPU predict.  Weights use the
reference's checkpoint naming (NNDynamicsModel/dense{,_1,_2}/{kernel,bias}:0, logz.py:78-85)."""
import numpy as np

from . import _lib
from .planner import Planner


class NNDynamicsModel:
    def __init__(self, env, n_layers, size, activation="relu", output_activation=None, normalization=None,
                 batch_size=None, iterations=None, learning_rate=None, sess=None, precision="fp32", layer_norm=False, device=None):
        if output_activation is not None:
            raise NotImplementedError("output_activation must be None (train_mpc_ppo.py:540)")
        self.env = env
        self.obs_dim = env.observation_space.shape[0]
        self.act_dim = env.action_space.shape[0]
        self.n_layers, self.size = n_layers, size
        self.activation = getattr(activation, "__name__", activation)  # accepts "relu"/"tanh" or a callable
        self.batch_size, self.iterations, self.learning_rate, self.sess = batch_size, iterations, learning_rate, sess
        self.scope = "NNDynamicsModel"
        if normalization is not None:
            (self.mean_obs, self.std_obs, self.mean_action, self.std_action, self.mean_reward, self.std_reward,
             self.mean_nxt_state, self.std_nxt_state, self.mean_deltas, self.std_deltas) = normalization  # dynamics.py:41
        self.weights = None
        self.device_weights = None     # (tensors, layout, stream_ptr) after set_weights_device
        self.layer_norm = layer_norm           # FLAGS.LAYER_NORM of the reference (dynamics.py:68-69)
        self.layer_norm_params = None          # [(gamma, beta)] per hidden layer
        self.version = 0           # bumped on every weight change; controllers re-upload lazily
        self.precision = precision
        self.device = device       # GPU of the stand-alone predict(); None = the tensors' device after set_weights_device, else the current device
        self._predictor = None

    def _predict_device(self):
        """Device of the helper planner behind predict(): the explicit one, else where device-resident weights live, else
        the calling thread's current CUDA device (never a hard-wired 0: on rank r of a multi-GPU job that would create a
        context on GPU 0 and dereference GPU-r pointers there)."""
        if self.device is not None:
            return int(self.device)
        if self.device_weights is not None:
            dev = getattr(self.device_weights[0][0][0], "device", None)
            if getattr(dev, "index", None) is not None:
                return int(dev.index)
        return _lib.current_device()

    # -- weight hand-off (replaces tf variables / fit) ------------------------------------------------
    def set_weights(self, weights, layer_norm_params=None):
        self.weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in weights]
        self.device_weights = None
        if self.layer_norm:
            if layer_norm_params is None:
                raise ValueError("this model was built with layer_norm=True: pass [(gamma, beta), ...] per hidden layer")
            self.layer_norm_params = [(np.asarray(g, np.float32), np.asarray(b, np.float32)) for g, b in layer_norm_params]
        self.version += 1
        self._predictor = None

    def set_weights_device(self, tensors, layout="out_in", stream_ptr=None, layer_norm_tensors=None):
        """Hand-off from a trainer that keeps the parameters on the planner's GPU (what replaces sess.run + fit when
        the dynamics model is trained in PyTorch): tensors = [(W, b), ...] float32 CUDA tensors, layout "out_in"
        (torch.nn.Linear.weight) or "in_out" (TF kernels); layer_norm_tensors = [(gamma, beta), ...] CUDA tensors when
        the model has LayerNorm.  No host copy is made; controllers and predict() pick the tensors up through
        Planner.set_dynamics_device / set_layer_norm_device on stream_ptr."""
        if self.layer_norm and layer_norm_tensors is None:
            raise ValueError("this model was built with layer_norm=True: pass layer_norm_tensors=[(gamma, beta), ...]")
        self.device_layer_norm = list(layer_norm_tensors) if self.layer_norm else None
        self.device_weights = (list(tensors), layout, stream_ptr)
        self.weights = None
        self.version += 1
        self._predictor = None

    def upload(self, planner):
        """Gives `planner` the current dynamics weights (and LayerNorm parameters)."""
        if self.device_weights is not None:
            tensors, layout, stream_ptr = self.device_weights
            planner.set_dynamics_device(tensors, layout=layout, stream_ptr=stream_ptr)
            if self.layer_norm:
                planner.set_layer_norm_device(self.device_layer_norm, stream_ptr=stream_ptr)
            return
        if self.weights is None:
            raise RuntimeError("NNDynamicsModel has no weights (set_weights / set_weights_device / load_vars)")
        planner.set_dynamics(self.weights)
        if self.layer_norm:
            planner.set_layer_norm(self.layer_norm_params)

    def load_vars(self, var_dict):
        """var_dict: {tf variable name: ndarray} as written by logz.pickle_tf_vars (logz.py:78-85)."""
        names = ["dense"] + [f"dense_{i}" for i in range(1, self.n_layers + 1)]
        ln = None
        if self.layer_norm:   # tf.contrib.layers.layer_norm variable names [from memory - not in the reference tree]
            lnn = ["LayerNorm"] + [f"LayerNorm_{i}" for i in range(1, self.n_layers)]
            ln = [(var_dict[f"{self.scope}/{n}/gamma:0"], var_dict[f"{self.scope}/{n}/beta:0"]) for n in lnn]
        self.set_weights([(var_dict[f"{self.scope}/{n}/kernel:0"], var_dict[f"{self.scope}/{n}/bias:0"])
                          for n in names], ln)

    def normalization6(self):
        return (self.mean_obs, self.std_obs, self.mean_action, self.std_action, self.mean_deltas, self.std_deltas)

    def fit(self, data):
        """dynamics.py:81-104 trains the TF graph.  This package is the INFERENCE side of the planning path only (SURVEY.md
        section 8: training is out of scope): train in the host framework and hand the new weights over with
        set_weights() / set_weights_device() after each fit."""
        raise NotImplementedError("bc_mpc_b200 is inference-only: train in the host framework, then call set_weights() / "
                                  "set_weights_device() (see INTEGRATION.md)")

    # -- dynamics.py:106-119 --------------------------------------------------------------------------
    def predict(self, unnormalized_state, unnormalized_action):
        s = np.asarray(unnormalized_state, np.float64)
        a = np.asarray(unnormalized_action, np.float64)
        if self._predictor is None:
            p = Planner(self.obs_dim, self.act_dim, self.n_layers, self.size, self.activation, num_paths=1,
                        horizon=1, precision="fp32", device=self._predict_device())
            self.upload(p)
            p.set_normalization(*self.normalization6())
            self._predictor = p
        return self._predictor.predict(s, a)


class NNDynamicsRewardModel(NNDynamicsModel):
    """Drop-in for the inference side of NNDynamicsRewardModel (dynamics.py:121-238): a shared tanh layer of 500 units, a
    tanh + linear head for the normalised state delta and a tanh + linear head for the normalised reward.  ``predict``
    returns ``(next_states, rewards[n,1])`` (dynamics.py:221-238).  The reference hard-codes size 500 and tanh
    (build_network, :162); both are parameters here with those defaults."""

    def __init__(self, env, normalization, batch_size=None, iterations=None, learning_rate=None, sess=None, size=500,
                 activation="tanh", layer_norm=False, device=None):
        super().__init__(env, n_layers=2, size=size, activation=activation, normalization=normalization,
                         batch_size=batch_size, iterations=iterations, learning_rate=learning_rate, sess=sess,
                         layer_norm=layer_norm, device=device)
        self.reward_weights = None
        self.reward_layer_norm = None

    def set_weights(self, weights, reward_weights, layer_norm_params=None, reward_layer_norm=None):
        """weights: share, delta-hidden, delta-out; reward_weights: reward-hidden, reward-out (TF creation order:
        dense, dense_1, dense_2, dense_3, dense_4 under scope NNDynamicsModel).  With layer_norm=True:
        layer_norm_params = [(gamma, beta)] of share and delta-hidden, reward_layer_norm = (gamma, beta) of reward-hidden
        (dynamics.py:165-176)."""
        self.reward_weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in reward_weights]
        if self.layer_norm:
            if reward_layer_norm is None:
                raise ValueError("this model was built with layer_norm=True: pass reward_layer_norm=(gamma, beta)")
            self.reward_layer_norm = tuple(np.asarray(t, np.float32) for t in reward_layer_norm)
        super().set_weights(weights, layer_norm_params)

    def load_vars(self, var_dict):
        names = ["dense", "dense_1", "dense_2", "dense_3", "dense_4"]
        ws = [(var_dict[f"{self.scope}/{n}/kernel:0"], var_dict[f"{self.scope}/{n}/bias:0"]) for n in names]
        ln = None
        if self.layer_norm:   # creation order: share, delta-hidden, reward-hidden (dynamics.py:165-176) [names from memory]
            ln = [(var_dict[f"{self.scope}/{n}/gamma:0"], var_dict[f"{self.scope}/{n}/beta:0"])
                  for n in ("LayerNorm", "LayerNorm_1", "LayerNorm_2")]
        self.set_weights(ws[:3], ws[3:], ln[:2] if ln else None, ln[2] if ln else None)

    def set_weights_device(self, tensors, reward_tensors, layout="out_in", stream_ptr=None, layer_norm_tensors=None,
                           reward_layer_norm_tensors=None):
        """Device hand-off of the whole reward model: tensors = share / delta-hidden / delta-out, reward_tensors =
        reward-hidden / reward-out [(W, b)] float32 CUDA tensors; with layer_norm=True also the (gamma, beta) tensors of
        share + delta-hidden (layer_norm_tensors) and of reward-hidden (reward_layer_norm_tensors)."""
        if self.layer_norm and reward_layer_norm_tensors is None:
            raise ValueError("this model was built with layer_norm=True: pass reward_layer_norm_tensors=(gamma, beta)")
        self.device_reward = (list(reward_tensors), reward_layer_norm_tensors if self.layer_norm else None)
        self.reward_weights = None
        super().set_weights_device(tensors, layout=layout, stream_ptr=stream_ptr, layer_norm_tensors=layer_norm_tensors)

    def upload_reward_head(self, planner, gamma=1.0):
        if self.device_weights is not None:
            _, layout, stream_ptr = self.device_weights
            rt, rln = self.device_reward
            planner.set_reward_head_device(rt, float(np.asarray(self.mean_reward)), float(np.asarray(self.std_reward)), gamma,
                                           layer_norm=rln, layout=layout, stream_ptr=stream_ptr)
            return
        planner.set_reward_head(self.reward_weights, float(np.asarray(self.mean_reward)), float(np.asarray(self.std_reward)), gamma,
                                layer_norm=self.reward_layer_norm if self.layer_norm else None)

    def predict(self, unnormalized_state, unnormalized_action):
        if self.weights is None and self.device_weights is None:
            raise RuntimeError("NNDynamicsRewardModel has no weights (set_weights / set_weights_device / load_vars)")
        if self._predictor is None:
            p = Planner(self.obs_dim, self.act_dim, 2, self.size, self.activation, num_paths=1, horizon=1, precision="fp32",
                        device=self._predict_device())
            self.upload(p)
            p.set_normalization(*self.normalization6())
            self.upload_reward_head(p)
            self._predictor = p
        return self._predictor.predict_reward(np.asarray(unnormalized_state, np.float64), np.asarray(unnormalized_action, np.float64))
