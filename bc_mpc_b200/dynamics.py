"""Drop-in for NNDynamicsModel's inference side (dynamics.py:8-52,106-119).

The reference builds a TF1 graph and trains it with `fit`; training is out of scope for this path
(SURVEY.md section 8a, row a5), so the model here is a weight container + GPU predict.  Weights use the
reference's checkpoint naming (NNDynamicsModel/dense{,_1,_2}/{kernel,bias}:0, logz.py:78-85)."""
import numpy as np

from .planner import Planner


class NNDynamicsModel:
    def __init__(self, env, n_layers, size, activation="relu", output_activation=None, normalization=None,
                 batch_size=None, iterations=None, learning_rate=None, sess=None, precision="fp32", layer_norm=False):
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
        self.layer_norm = layer_norm           # FLAGS.LAYER_NORM of the reference (dynamics.py:68-69)
        self.layer_norm_params = None          # [(gamma, beta)] per hidden layer
        self.version = 0           # bumped on every weight change; controllers re-upload lazily
        self.precision = precision
        self._predictor = None

    # -- weight hand-off (replaces tf variables / fit) ------------------------------------------------
    def set_weights(self, weights, layer_norm_params=None):
        self.weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in weights]
        if self.layer_norm:
            if layer_norm_params is None:
                raise ValueError("this model was built with layer_norm=True: pass [(gamma, beta), ...] per hidden layer")
            self.layer_norm_params = [(np.asarray(g, np.float32), np.asarray(b, np.float32)) for g, b in layer_norm_params]
        self.version += 1
        self._predictor = None

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
        raise NotImplementedError("training stays in the host framework; call set_weights() after each fit")

    # -- dynamics.py:106-119 --------------------------------------------------------------------------
    def predict(self, unnormalized_state, unnormalized_action):
        if self.weights is None:
            raise RuntimeError("NNDynamicsModel has no weights (set_weights / load_vars)")
        s = np.asarray(unnormalized_state, np.float64)
        a = np.asarray(unnormalized_action, np.float64)
        if self._predictor is None:
            p = Planner(self.obs_dim, self.act_dim, self.n_layers, self.size, self.activation, num_paths=1,
                        horizon=1, precision="fp32")
            p.set_dynamics(self.weights)
            if self.layer_norm:
                p.set_layer_norm(self.layer_norm_params)
            p.set_normalization(*self.normalization6())
            self._predictor = p
        return self._predictor.predict(s, a)
