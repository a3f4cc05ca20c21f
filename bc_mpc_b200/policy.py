"""Inference side of the proposal networks: MlpPolicy.act (ppo_bc_policy.py:56-88,174-185) and
BCnetwork.predict (behavioral_cloning.py:21-26,44-47).  Weight containers + GPU forward."""
import numpy as np

from . import _lib
from .planner import Planner


def rms_from_sums(runningsum, runningsumsq, count):
    """baselines RunningMeanStd mean / std (float32 casts; std = sqrt(max(E[x^2]-mean^2, 1e-2)))."""
    mean = (np.asarray(runningsum, np.float64) / count).astype(np.float32)
    msq = (np.asarray(runningsumsq, np.float64) / count).astype(np.float32)
    return mean, np.sqrt(np.maximum(msq - mean * mean, np.float32(1e-2))).astype(np.float32)


class _ProposalNet:
    logstd = None
    filt_mean = None
    filt_std = None
    clip = 0.0

    def __init__(self, env, device=None):
        self.ob_dim = env.observation_space.shape[0]
        self.ac_dim = env.action_space.shape[0]
        self.weights = None
        self.device_weights = None      # (tensors, layout, stream_ptr) after set_weights_device
        self.version = 0
        self.device = device            # GPU of the stand-alone forward; None = the tensors' device, else the current device
        self._fwd = None

    def _forward_device(self):
        if self.device is not None:
            return int(self.device)
        if self.device_weights is not None:
            dev = getattr(self.device_weights[0][0][0], "device", None)
            if getattr(dev, "index", None) is not None:
                return int(dev.index)
        return _lib.current_device()

    def set_weights_device(self, tensors, logstd=None, filt_mean=None, filt_std=None, layout="out_in", stream_ptr=None):
        """Hand-off from a trainer that keeps the policy parameters on the planner's GPU (PPO / BC updates in PyTorch):
        tensors = [(W, b), ...] float32 CUDA tensors ("out_in" = torch.nn.Linear.weight, "in_out" = TF kernels); logstd
        and the observation-filter statistics are small host arrays."""
        self.device_weights = (list(tensors), layout, stream_ptr)
        self.weights = None
        if logstd is not None:
            self.logstd = np.asarray(logstd, np.float32).reshape(-1)
        if filt_mean is not None:
            self.filt_mean, self.filt_std = np.asarray(filt_mean, np.float32), np.asarray(filt_std, np.float32)
        self.version += 1
        self._fwd = None

    def set_weights(self, weights, logstd=None, filt_mean=None, filt_std=None):
        self.device_weights = None
        self.weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in weights]
        if logstd is not None:
            self.logstd = np.asarray(logstd, np.float32).reshape(-1)
        if filt_mean is not None:
            self.filt_mean, self.filt_std = np.asarray(filt_mean, np.float32), np.asarray(filt_std, np.float32)
        self.version += 1
        self._fwd = None

    def upload(self, planner):
        if self.device_weights is not None:
            tensors, layout, stream_ptr = self.device_weights
            planner.set_policy_device(tensors, self.logstd, self.filt_mean, self.filt_std, self.clip, layout=layout, stream_ptr=stream_ptr)
            return
        planner.set_policy(self.weights, self.logstd, self.filt_mean, self.filt_std, self.clip)

    def _forward(self, ob, noise=None):
        if self.weights is None and self.device_weights is None:
            raise RuntimeError("proposal network has no weights (set_weights / set_weights_device / load_vars)")
        if self._fwd is None:
            p = Planner(self.ob_dim, self.ac_dim, 1, 4, num_paths=1, horizon=1, device=self._forward_device())
            self.upload(p)
            self._fwd = p
        return self._fwd.policy_act(ob, noise)


class MlpPolicy(_ProposalNet):
    """ppo_bc_policy.MlpPolicy, inference side: the 'pol' branch (action mean / sample) and the 'vf' branch (value)."""
    recurrent = False
    clip = 5.0

    def __init__(self, sess=None, env=None, hid_size=128, num_hid_layers=2, clip_param=None, entcoeff=None, device=None):
        super().__init__(env, device=device)
        self.hid_size, self.num_hid_layers = hid_size, num_hid_layers
        self.vf_weights = None          # [(kernel, bias)] of pi/vf/fc*/final (ppo_bc_policy.py:59-64)
        self._vf = None

    def set_value_weights(self, weights):
        """Value branch (ppo_bc_policy.py:59-64): tanh layers fc1..fcN of hid_size units, linear `final` -> 1."""
        self.vf_weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in weights]
        if self.vf_weights[-1][0].shape[1] != 1:
            raise ValueError("the value head ends in one unit")
        self._vf = None

    def load_vars(self, var_dict, scope="pi/pi"):
        names = [f"fc{i + 1}" for i in range(self.num_hid_layers)] + ["final"]
        ws = [(var_dict[f"{scope}/pol/{n}/kernel:0"], var_dict[f"{scope}/pol/{n}/bias:0"]) for n in names]
        mean, std = rms_from_sums(var_dict[f"{scope}/obfilter/runningsum:0"],
                                  var_dict[f"{scope}/obfilter/runningsumsq:0"],
                                  float(var_dict[f"{scope}/obfilter/count:0"]))
        self.set_weights(ws, var_dict[f"{scope}/pol/logstd:0"], mean, std)
        if f"{scope}/vf/final/kernel:0" in var_dict:
            self.set_value_weights([(var_dict[f"{scope}/vf/{n}/kernel:0"], var_dict[f"{scope}/vf/{n}/bias:0"]) for n in names])

    def set_weights(self, weights, logstd=None, filt_mean=None, filt_std=None):
        super().set_weights(weights, logstd, filt_mean, filt_std)
        self._vf = None                 # the value head shares the observation filter

    def value(self, ob):
        """vpred of ppo_bc_policy.py:64: the 'vf' MLP on the same filtered, clipped observation; float32 [n].  Evaluated by
        the same device kernel as the action mean (a proposal-network forward with one output unit)."""
        if self.vf_weights is None:
            return None
        ob = np.asarray(ob, np.float64)
        if ob.ndim == 1:
            ob = ob[None]
        if self._vf is None:
            p = Planner(self.ob_dim, 1, 1, 4, num_paths=1, horizon=1, device=self._forward_device())
            p.set_policy(self.vf_weights, None, self.filt_mean, self.filt_std, self.clip)
            self._vf = p
        return self._vf.policy_act(ob)[:, 0]

    def act(self, ob, stochastic=True):
        """ppo_bc_policy.py:174-185: returns (ac [n,A], vpred [n]).  vpred is None when no value weights were given
        (set_value_weights / load_vars with pi/vf/*): the planners never use it (controllers.py:203 discards it), the PPO
        segment generator does (utils.py:196,238) and says so."""
        ob = np.asarray(ob, np.float64)
        if ob.ndim == 1:
            ob = ob[None]
        noise = np.random.standard_normal((ob.shape[0], self.ac_dim)).astype(np.float32) if stochastic else None
        return self._forward(ob, noise), self.value(ob)


class BCnetwork(_ProposalNet):
    """behavioral_cloning.BCnetwork: 20 -> 128 tanh -> 64 tanh -> A linear, no observation filter."""

    def __init__(self, sess=None, env=None, batch_size=None, learning_rate=None, device=None):
        super().__init__(env, device=device)

    def predict(self, obs):
        return self._forward(np.asarray(obs, np.float64))
