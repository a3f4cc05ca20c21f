#!/usr/bin/env python
"""bench.py - throughput of the MPC planning hot path (BASELINE.json metric: dynamics transitions/s and
plans/s, HalfCheetah).

Workload (config.workload): the per-GPU share of BASELINE.json configs[2] - 32*N concurrent synthetic root
states x K=4096 paths x H=20, MLP 26->500->500->20 ReLU, cheetah cost, uniform shooting sampled on the
device, bf16 fast mode; the K paths of every state are sharded over the N ranks (512*8/N ... K/N per rank)
and each plan ends with one exchange of 80-byte records (NVLink peer stores; NCCL all-gather + min-loc kernel as fallback).
N=8 is configs[2] itself.
A "step" is one batched plan.  `value` = transitions/s with inputs resident in HBM; `e2e` = the same
through the host-buffer C-ABI call (pinned-host -> device copy of the states and device -> host read of the
winning records inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # The reference arm is a host-core measurement: use every core.  torch.distributed.run exports OMP_NUM_THREADS=1 to its
    # workers when nproc > 1, and OpenBLAS sizes its pool from the environment when numpy is first imported - so the
    # thread count is fixed here, before that import, and reported from threadpoolctl in the line.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN = (500, 500)
K_PATHS, HORIZON, STATES_PER_GPU = 4096, 20, 32
FLOP_PER_TRANSITION = 2 * (26 * 500 + 500 * 500 + 500 * 20)   # 546 000; padding not credited
# dram__bytes_read.sum + dram__bytes_write.sum of one rollout launch of this workload, from the committed ncu --set full
# capture (the kernel streams its 0.6 MB of weights from L2; states in, 80-byte records out)
NCU_DRAM_BYTES_PER_LAUNCH = 697088     # dram__bytes_read.sum + dram__bytes_write.sum of rollout_tmem_kernel, one ncu --set full capture
NCU_TRAFFIC_SOURCE = "profiles/r2zm_ncu_full_rollout_tmem.csv"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6550.0   # fallback: B200_PROFILING.md


_NVML_POLLER = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
        nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
while True:
    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, *[int(bool(r & b)) for b in bits],
          nv.nvmlDeviceGetPowerUsage(h) / 1000.0, flush=True)
    time.sleep(0.01)
"""


class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions (B200_PROFILING.md recipe).  A helper PROCESS polls NVML
    every ~10 ms (faster polling measurably slows the host-buffer calls: NVML and the CUDA driver share locks) (nvidia_ml_py: the counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; a thread in
    this process would take the GIL from the timed host-buffer calls); `nvidia-smi -lms` is the fallback.  Only samples
    that arrive inside a marked region count."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.regions, self._t0, self.source = index, [], None, [], None, None
        self.log = None        # the poller writes to a file that is parsed after the run: nothing of it runs in this process

    def start(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
        try:
            import pynvml  # noqa: F401
            import tempfile
            self.log = tempfile.NamedTemporaryFile("w+", prefix="bcmpc_clocks_", suffix=".txt", delete=False)
            self.proc = subprocess.Popen([sys.executable, "-c", _NVML_POLLER, str(phys)], stdout=self.log,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvml"
        except Exception:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={phys}", f"--query-gpu={self.Q}",
                                              "--format=csv,noheader,nounits", "-lms", "20"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.source = "nvidia-smi"
            except Exception:
                self.proc = None
        if self.proc is not None and self.source == "nvidia-smi":
            threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_ready(self, timeout=5.0):
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < timeout:
            if self.rows or (self.log is not None and os.path.getsize(self.log.name) > 0):
                return
            time.sleep(0.01)

    def begin(self):
        self._t0 = time.time()

    def end(self):
        self.regions.append((self._t0, time.time()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        if self.log is not None:
            self.proc.wait()
            for line in open(self.log.name):
                f = line.split()
                if len(f) == 8:
                    self.rows.append((float(f[0]), [f[1], f[2]] + ["Active" if x == "1" else "Not Active" for x in f[3:7]] + [f[7]]))
            os.unlink(self.log.name)
        inside = [r for t, r in self.rows if any(a <= t <= b for a, b in self.regions)]
        rows = inside if inside else [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = sorted({self.NAMES[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        pw = [float(r[6]) for r in rows if len(r) >= 7]
        cap = [r for r in rows if len(r) >= 6 and r[5].lower() == "active"]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "samples_inside_timed_regions": len(inside), "source": self.source,
                "power_w_median": float(np.median(pw)) if pw else None, "power_w_max": max(pw) if pw else None,
                "sw_power_cap_share": len(cap) / len(rows) if rows else None}


def problem():
    from oracle import mpc_oracle as O   # synthetic inputs only (shared generator); never on the timed path
    rng, norm, weights, state = O.synthetic_problem(0, HIDDEN)
    return O, rng, norm, weights


def cpu_baseline(budget_s=12.0):
    """The oracle port (reference controllers/cost restated + numpy float32 predict) on the host cores:
    BASELINE.json configs[0] (1 state, K=1000, H=15) repeated for about `budget_s` seconds."""
    O, rng, norm, weights = problem()
    dyn = O.OracleDynamics(weights, norm)
    K, H = 1000, 15
    mpc = O.OracleMPC(dyn, H, K)
    state = O.synthetic_state(rng, norm)
    mpc.get_action(state, rng.uniform(-1, 1, (H, K, 6)))
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s or n < 3:
        mpc.get_action(state, rng.uniform(-1, 1, (H, K, 6)))   # controllers.py:53 sampling included
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * K * H / dt, "unit": "transitions/s", "cores": blas_threads(), "kind": "port",
            "sample": f"{n} plans of configs[0] (1 state x K=1000 x H=15, numpy float32 BLAS predict, {blas_threads()} BLAS threads on {os.cpu_count()} host cores), "
                      f"{dt:.1f} s; plans/s = {n / dt:.2f}; TF-less restatement = optimistic stand-in for the TF CPU path"}


def extras():
    """Not part of the contract line's headline: the other BASELINE.json configs on 1 GPU, a few calls each.
    configs[0]/[1] are single-plan latency cases (host-buffer C-ABI call, copies included); configs[3] is the
    policy-guided variant (proposal network + Gaussian noise fused with the rollout), 16 states per call."""
    from bc_mpc_b200 import Planner
    O, rng, norm, weights = problem()
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
    state = O.synthetic_state(rng, norm)
    out = {}

    def make(K, H, precision, proposal="uniform", B=1):
        pl = Planner(20, 6, 2, HIDDEN[0], "relu", num_paths=K, horizon=H, max_states=B, precision=precision, proposal=proposal)
        pl.set_dynamics(weights)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"],
                             norm["mean_deltas"], norm["std_deltas"])
        if proposal != "uniform":
            pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
        return pl

    def time_plans(pl, states, n=20):
        for i in range(3):
            pl.plan(states, seed=i)
        t0 = time.perf_counter()
        for i in range(n):
            pl.plan(states, seed=100 + i)
        return (time.perf_counter() - t0) / n

    for tag, K, H, prec in (("configs0_K1000_H15_bf16", 1000, 15, "bf16"), ("configs0_K1000_H15_fp32_verify", 1000, 15, "fp32"),
                            ("configs1_K4096_H15_fp32_verify", 4096, 15, "fp32"), ("configs1_K4096_H15_bf16", 4096, 15, "bf16")):
        pl = make(K, H, prec)
        dt = time_plans(pl, state)
        out[tag] = {"plan_latency_ms": 1e3 * dt, "plans_per_s": 1 / dt, "transitions_per_s": K * H / dt}
        pl.close()
    B, K, H = 16, 8192, 15
    pl = make(K, H, "bf16", "policy_gaussian", B)
    states = norm["mean_obs"] + norm["std_obs"] * np.random.default_rng(2).standard_normal((B, 20))
    states[:, 18] = 0.0
    pl.enable_timing(True)
    dt = time_plans(pl, states, 10)
    k_ms = pl.last_kernel_ms()
    flop = FLOP_PER_TRANSITION + 2 * (20 * 128 + 128 * 128 + 128 * 6)
    out["configs3_policy_gaussian_K8192_H15_bf16_x16states"] = {
        "transitions_per_s_e2e": B * K * H / dt, "transitions_per_s_kernel": B * K * H / (k_ms * 1e-3), "kernel_ms": k_ms,
        "tflops_algorithmic": B * K * H * flop / (k_ms * 1e-3) / 1e12, "flop_per_transition": flop}
    pl.close()
    # the same plan with MlpPolicy(hid_size=256) (various_train/train_mpc_ppo_r.py:142): the kVarPolWide kernel variant -
    # an 88 KB resident proposal network per CTA, quarter-chunk weight stages for the dynamics net
    pw_w, logstd_w, rms_mean_w, rms_std_w = O.synthetic_policy(np.random.default_rng(4), hidden=(256, 256), norm=norm)
    pl = make(K, H, "bf16", "policy_gaussian", B)
    pl.set_policy(pw_w, logstd_w, rms_mean_w, rms_std_w, clip=5.0)
    pl.enable_timing(True)
    dt = time_plans(pl, states, 10)
    k_ms = pl.last_kernel_ms()
    flop_w = FLOP_PER_TRANSITION + 2 * (20 * 256 + 256 * 256 + 256 * 6)
    out["configs3_policy_gaussian_hid256_K8192_H15_bf16_x16states"] = {
        "transitions_per_s_e2e": B * K * H / dt, "transitions_per_s_kernel": B * K * H / (k_ms * 1e-3), "kernel_ms": k_ms,
        "tflops_algorithmic": B * K * H * flop_w / (k_ms * 1e-3) / 1e12, "flop_per_transition": flop_w}
    pl.close()
    # the reference's default network size (train_mpc_ppo.py:75 size=256; the shape of saved_data/ppo-mpc/vars.pkl):
    # hidden <= 256 runs on rollout_narrow_kernel (two tiles per CTA); 37 states x 32 tiles = 296 CTA-pair units = 4 even waves
    B, K, H = 37, 4096, 20
    rng2, norm2, w2, _ = O.synthetic_problem(0, (256, 256))
    pl = Planner(20, 6, 2, 256, "relu", num_paths=K, horizon=H, max_states=B, precision="bf16")
    pl.set_dynamics(w2)
    pl.set_normalization(norm2["mean_obs"], norm2["std_obs"], norm2["mean_action"], norm2["std_action"], norm2["mean_deltas"], norm2["std_deltas"])
    st2 = norm2["mean_obs"] + norm2["std_obs"] * np.random.default_rng(5).standard_normal((B, 20))
    pl.enable_timing(True)
    dt = time_plans(pl, st2, 10)
    k_ms = pl.last_kernel_ms()
    flop2 = 2 * (26 * 256 + 256 * 256 + 256 * 20)
    out["net_2x256_K4096_H20_bf16_x37states"] = {
        "transitions_per_s_e2e": B * K * H / dt, "transitions_per_s_kernel": B * K * H / (k_ms * 1e-3), "kernel_ms": k_ms,
        "tflops_algorithmic": B * K * H * flop2 / (k_ms * 1e-3) / 1e12, "flop_per_transition": flop2}
    pl.close()
    # configs[4] (K = 2**20 x H = 30, MLP 3 x 1024) on the float32 verification kernel, for comparison with the tensor-core line
    # `configs4` of the contract line (cluster-split tcgen05 kernel, rollout_deep.cuh): one plan, ~5 s
    rng4, norm4, w4, s4 = O.synthetic_problem(5, (1024, 1024, 1024))
    pl = Planner(20, 6, 3, 1024, "relu", num_paths=1 << 20, horizon=30, precision="fp32")
    pl.set_dynamics(w4)
    pl.set_normalization(norm4["mean_obs"], norm4["std_obs"], norm4["mean_action"], norm4["std_action"], norm4["mean_deltas"], norm4["std_deltas"])
    t0 = time.perf_counter()
    pl.plan(s4, seed=1)
    dt = time.perf_counter() - t0
    out["configs4_K1048576_H30_3x1024_fp32_verify"] = {"plan_s": dt, "transitions_per_s": (1 << 20) * 30 / dt,
                                                       "tflops_fp32": (1 << 20) * 30 * 4288512 / dt / 1e12}
    pl.close()
    # ... and the same net under policy-guided shooting (train_mpc_ppo.py --n_layers 3 --size 1024 with its MlpPolicy(hid_size=128)):
    # the proposal network resident in every CTA of the cluster-split kernel
    pl = Planner(20, 6, 3, 1024, "relu", num_paths=1 << 20, horizon=30, precision="bf16", proposal="policy_gaussian")
    pl.set_dynamics(w4)
    pl.set_normalization(norm4["mean_obs"], norm4["std_obs"], norm4["mean_action"], norm4["std_action"], norm4["mean_deltas"], norm4["std_deltas"])
    pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    pl.enable_timing(True)
    dt = time_plans(pl, s4, 4)
    k_ms = pl.last_kernel_ms()
    flop4 = 4288512 + 2 * (20 * 128 + 128 * 128 + 128 * 6)
    out["net_3x1024_policy_gaussian_K1048576_H30_bf16"] = {
        "plan_ms": 1e3 * dt, "kernel_ms": k_ms, "transitions_per_s_e2e": (1 << 20) * 30 / dt,
        "tflops_algorithmic": (1 << 20) * 30 * flop4 / (k_ms * 1e-3) / 1e12, "flop_per_transition": flop4}
    pl.close()
    del w4
    # learned-reward planner (NNDynamicsRewardModel + MPCcontrollerReward: what run.sh launches), tanh, 26-500-{500-20, 500-1}
    B, K, H = 32, 4096, 20
    rrng, rnorm, rw, rrw, _ = O.synthetic_reward_problem(3, 500)
    pl = Planner(20, 6, 2, 500, "tanh", num_paths=K, horizon=H, max_states=B, precision="bf16")
    pl.set_dynamics(rw)
    pl.set_normalization(rnorm["mean_obs"], rnorm["std_obs"], rnorm["mean_action"], rnorm["std_action"], rnorm["mean_deltas"], rnorm["std_deltas"])
    pl.set_reward_head(rrw, rnorm["mean_reward"], rnorm["std_reward"], 0.99)
    states = rnorm["mean_obs"] + rnorm["std_obs"] * np.random.default_rng(2).standard_normal((B, 20))
    pl.enable_timing(True)
    dt = time_plans(pl, states, 10)
    k_ms = pl.last_kernel_ms()
    flop = 2 * (26 * 500 + 2 * 500 * 500 + 500 * 20 + 500)
    out["learned_reward_tanh_K4096_H20_bf16_x32states"] = {
        "transitions_per_s_e2e": B * K * H / dt, "transitions_per_s_kernel": B * K * H / (k_ms * 1e-3), "kernel_ms": k_ms,
        "tflops_algorithmic": B * K * H * flop / (k_ms * 1e-3) / 1e12, "flop_per_transition": flop}
    pl.close()
    # the configuration run.sh launches: LEARN_REWARD + LAYER_NORM (default) + SELFEXP (stochastic policy proposals),
    # mpc_horizon=30, simulated_paths=400; one env (the reference's loop) and 256 envs in lock step (sample_vectorized)
    lns = [(np.ones(500, np.float32), np.zeros(500, np.float32))] * 3
    for tag, B in (("runsh_default_1env", 1), ("runsh_default_256envs", 256)):
        K, H = 400, 30
        pl = Planner(20, 6, 2, 500, "tanh", num_paths=K, horizon=H, max_states=B, precision="bf16", proposal="policy_gaussian")
        pl.set_dynamics(rw)
        pl.set_layer_norm(lns[:2])
        pl.set_normalization(rnorm["mean_obs"], rnorm["std_obs"], rnorm["mean_action"], rnorm["std_action"], rnorm["mean_deltas"], rnorm["std_deltas"])
        pl.set_reward_head(rrw, rnorm["mean_reward"], rnorm["std_reward"], 1.0, layer_norm=lns[2])
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
        st = rnorm["mean_obs"] + rnorm["std_obs"] * np.random.default_rng(3).standard_normal((B, 20))
        pl.enable_timing(True)
        dt = time_plans(pl, st, 10)
        k_ms = pl.last_kernel_ms()
        out[tag + "_reward_layernorm_selfexp_K400_H30_bf16"] = {"plan_call_ms": 1e3 * dt, "plans_per_s": B / dt, "transitions_per_s_e2e": B * K * H / dt,
                                                               "kernel_ms": k_ms}
        pl.close()
    # train_mpc_ppo.py's own flag defaults: NNDynamicsModel size=256 with LAYER_NORM, MPCcontrollerPolicyNet with SELFEXP=False,
    # MPC_EXP=0.5 (deterministic policy mixed with uniform exploration), K=400, H=7 - on the narrow-net kernel (hidden <= 256: two tiles per CTA); one env and 256 envs in lock step
    rng5, norm5, w5, _ = O.synthetic_problem(0, (256, 256))
    ln5 = [(np.ones(256, np.float32), np.zeros(256, np.float32))] * 2
    for tag, B in (("flag_defaults_1env", 1), ("flag_defaults_256envs", 256)):
        K, H = 400, 7
        pl = Planner(20, 6, 2, 256, "relu", num_paths=K, horizon=H, max_states=B, precision="bf16", proposal="policy_mix", explore=0.5)
        pl.set_dynamics(w5)
        pl.set_layer_norm(ln5)
        pl.set_normalization(norm5["mean_obs"], norm5["std_obs"], norm5["mean_action"], norm5["std_action"], norm5["mean_deltas"], norm5["std_deltas"])
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
        st = norm5["mean_obs"] + norm5["std_obs"] * np.random.default_rng(3).standard_normal((B, 20))
        pl.enable_timing(True)
        dt = time_plans(pl, st, 20)
        out[tag + "_2x256_layernorm_policy_mix_K400_H7_bf16"] = {"plan_call_ms": 1e3 * dt, "plans_per_s": B / dt, "transitions_per_s_e2e": B * K * H / dt,
                                                              "kernel_ms": pl.last_kernel_ms()}
        pl.close()
    # stand-alone sampling stage (replay / debugging aid; the planner itself draws in registers): HBM-bound writes,
    # 6 float64 per (state, step, path) = 48 B per transition
    pl = make(1 << 17, 20, "bf16", B=4)
    pl.enable_timing(True)
    for i in range(3):
        inj = pl.sample_injection(7 + i, n_states=4)
    ms = pl.last_kernel_ms()
    out["sampling_stage_standalone"] = {"bytes_written": int(inj.nbytes), "kernel_ms": ms, "GB_per_s": inj.nbytes / (ms * 1e-3) / 1e9,
                                        "hbm_peak_GB_per_s": hbm_peak(), "transitions_per_s": inj.size / 6 / (ms * 1e-3)}
    del inj
    # stand-alone cost stage (cheetah_cost_fn over given [n,20] float64 states / next states; in the planner it is fused)
    n = 1 << 21
    S0 = np.random.default_rng(4).standard_normal((n, 20))
    for _ in range(2):
        pl.cheetah_cost(S0, S0 + 0.01)
    ms = pl.last_kernel_ms()
    out["cost_stage_standalone"] = {"rows": n, "kernel_ms": ms, "GB_per_s_arrays": (2 * S0.nbytes + 8 * n) / (ms * 1e-3) / 1e9,
                                    "GB_per_s_touched_sectors": (3 * 32 + 8) * n / (ms * 1e-3) / 1e9, "hbm_peak_GB_per_s": hbm_peak(),
                                    "note": "reads 5 of 40 doubles per row (3 of 10 32-byte sectors), but ncu shows DRAM fetching 604 of the 671 MB of arrays at 6.6 TB/s (profiles/r1s_ncu_cost_stage.csv): HBM-bound for the row-major layout"}
    del S0
    pl.close()
    # weight refresh after dyn_model.fit (train_mpc_ppo.py:279): host arrays vs device-resident torch parameters
    import torch
    pl = make(4096, 15, "bf16")
    tens = [(torch.from_numpy(np.ascontiguousarray(w.T)).cuda(), torch.from_numpy(b).cuda()) for w, b in weights]
    torch.cuda.synchronize()
    for _ in range(3):
        pl.set_dynamics(weights)
        pl.set_dynamics_device(tens, layout="out_in", stream_ptr=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        pl.set_dynamics(weights)
    t_host = (time.perf_counter() - t0) / n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        pl.set_dynamics_device(tens, layout="out_in", stream_ptr=torch.cuda.current_stream().cuda_stream)
    e1.record()
    t_enq = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    out["weight_refresh_bf16_2x500"] = {"host_arrays_ms": 1e3 * t_host, "device_tensors_enqueue_ms": 1e3 * t_enq,
                                        "device_tensors_gpu_ms": e0.elapsed_time(e1) / n}
    pl.close()
    return out


def blas_threads():
    """Threads the BLAS behind numpy will actually use (threadpoolctl), not the core count."""
    try:
        from threadpoolctl import threadpool_info
        n = [int(d["num_threads"]) for d in threadpool_info() if d.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return 1


CONFIGS4_K, CONFIGS4_H, CONFIGS4_HIDDEN = 1 << 20, 30, (1024, 1024, 1024)
CONFIGS4_FLOP = 2 * (26 * 1024 + 2 * 1024 * 1024 + 1024 * 20)      # 4 288 512 per transition


def configs4_leg(torch, dist, world, rank, local, n_plans=6):
    """BASELINE.json configs[4] - the scaling-sweep config: ONE root state, K = 2**20 paths x H = 30, dynamics MLP 3 x 1024,
    K sharded over the ranks (strong scaling), NCCL all-gather of one 80-byte record per rank + min-loc.  Runs on the
    cluster-split tcgen05 kernel (rollout_deep.cuh).  Device time per plan from CUDA events, max over ranks."""
    from bc_mpc_b200.distributed import ShardedPlanner
    from oracle import mpc_oracle as O      # synthetic inputs only
    rng, norm, weights, state = O.synthetic_problem(5, CONFIGS4_HIDDEN)
    sp = ShardedPlanner(CONFIGS4_K, CONFIGS4_H, rank=rank, world_size=world, device=local, obs_dim=20, act_dim=6, n_hidden=3,
                        hidden_size=1024, activation="relu", max_states=1, precision="bf16", proposal="uniform")
    pl = sp.planner
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.set_action_bounds(-np.ones(6), np.ones(6))
    d_state = torch.from_numpy(state[None]).cuda()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(2):
        sp.plan_device(d_state, 1, seed=10 + i)
    barrier()
    pl.enable_timing(True)
    evs, kms = [], []
    for i in range(n_plans):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = sp.plan_device(d_state, 1, seed=100 + i)
        e1.record(stream)
        evs.append((e0, e1))
        torch.cuda.synchronize()
        kms.append(pl.last_kernel_ms())
    barrier()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs), sum(kms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    plan_ms, kernel_ms = float(ms[0].item()) / n_plans, float(ms[1].item()) / n_plans
    rec = np.frombuffer(out.cpu().numpy().tobytes(), dtype=[("cost", "<f8"), ("index", "<i8"), ("action", "<f8", (8,))])[0]
    pl.close()
    tr = CONFIGS4_K * CONFIGS4_H
    peak, _ = measured_peaks()
    tf_gpu = tr / world * CONFIGS4_FLOP / (kernel_ms * 1e-3) / 1e12
    return {"workload": f"BASELINE.json configs[4]: 1 root state x K=2^20 x H=30, MLP 26-1024-1024-1024-20 ReLU, bf16 fast mode, K sharded over {world} GPU(s)",
            "scaling": "strong", "n_gpus": world, "plans": n_plans, "plan_ms": plan_ms, "plans_per_s": 1e3 / plan_ms,
            "transitions_per_s": tr / (plan_ms * 1e-3), "kernel_ms_max_over_ranks": kernel_ms,
            "tflops_algorithmic_per_gpu": tf_gpu, "frac_of_burst_peak": tf_gpu / peak, "flop_per_transition": CONFIGS4_FLOP,
            "kernel": "bcmpc::dpx::rollout_deep_kernel<8>", "winner_index": int(rec["index"]), "winner_cost": float(rec["cost"])}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; TensorFlow is not installable here) on the host
    cores, each step a bounded sample of this arm's workload: 1 root state x K=4096 x H=20 (the workload is 32*N such
    states per step; the rate in transitions/s is comparable, the work per step is not - the line's config says so)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = blas_threads()
    O, rng, norm, weights = problem()
    dyn = O.OracleDynamics(weights, norm)
    mpc = O.OracleMPC(dyn, HORIZON, K_PATHS)
    state = O.synthetic_state(rng, norm)
    for _ in range(args.warmup):
        mpc.get_action(state, rng.uniform(-1, 1, (HORIZON, K_PATHS, 6)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mpc.get_action(state, rng.uniform(-1, 1, (HORIZON, K_PATHS, 6)))
    dt = time.perf_counter() - t0
    val = args.steps * K_PATHS * HORIZON / dt
    sample = (f"each step = 1 root state x K={K_PATHS} x H={HORIZON} of the workload ({K_PATHS * HORIZON} transitions), oracle port "
              f"(reference controllers.py / cost_functions.py logic + numpy float32 BLAS predict), {threads} BLAS threads on "
              f"{os.cpu_count()} host cores (OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')})")
    cfg = workload_config(args.gpus)
    cfg.update({"states": 1, "states_of_the_workload": STATES_PER_GPU * args.gpus,
                "precision": "float32 MLP / float64 normalise, state carry and cost (dynamics.py:106-119)",
                "parallelism": f"host CPU, {threads} BLAS threads", "l2": "n/a (host)",
                "sample": "1 of the workload's root states per step"})
    line = {"impl": "reference", "metric": "dynamics transitions/s (HalfCheetah random-shooting MPC)", "value": val,
            "unit": "transitions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "plans_per_s": args.steps / dt,
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": "transitions/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": f"BASELINE.json configs[2] per-GPU share: {STATES_PER_GPU * n} root states x K={K_PATHS} x H={HORIZON}, "
                        f"MLP 26-500-500-20 ReLU, cheetah_cost_fn, uniform shooting (device Philox), K sharded over {n} GPU(s)",
            "states": STATES_PER_GPU * n, "num_paths": K_PATHS, "horizon": HORIZON, "hidden": list(HIDDEN),
            "precision": "bf16 operands / fp32 accumulate, state and cost",
            "parallelism": f"paths sharded x{n}; per plan one 80 B record per state and rank, min-loc over NVLink peer stores (torch symmetric memory; NCCL all-gather fallback)" if n > 1 else "1 GPU",
            "l2": "flushed between steps (256 MiB memset outside the per-step event pairs); the 0.6 MB weight images are re-read from HBM once per step"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bc_mpc_b200 import _lib as L
    from bc_mpc_b200.distributed import ShardedPlanner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = STATES_PER_GPU * world
    O, rng, norm, weights = problem()
    sp = ShardedPlanner(K_PATHS, HORIZON, rank=rank, world_size=world, device=local, obs_dim=20, act_dim=6, n_hidden=2,
                        hidden_size=HIDDEN[0], activation="relu", max_states=B, precision="bf16", proposal="uniform")
    pl = sp.planner
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"],
                         norm["mean_deltas"], norm["std_deltas"])
    pl.set_action_bounds(-np.ones(6), np.ones(6))
    states = norm["mean_obs"] + norm["std_obs"] * np.random.default_rng(1).standard_normal((B, 20))
    states[:, 18] = 0.0
    d_states = torch.from_numpy(states).cuda()
    h_states = torch.from_numpy(states).pin_memory()
    h_outs = [torch.zeros(B * 80, dtype=torch.uint8).pin_memory() for _ in range(2)]
    e2e_done = [torch.cuda.Event(), torch.cuda.Event()]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        return sp.plan_device(d_states, B, seed=1000 + i)

    def step_e2e(i):
        if world == 1:
            # the reference-facing call itself: bcmpc_plan with HOST buffers (numpy in, numpy records out); the library
            # stages through its pinned buffers, copies host -> device, plans, copies the records back and synchronises
            return sp.planner.plan(states, seed=2000 + i)
        # every step: pinned host -> device copy of the states, plan (rollout + cross-rank exchange), device -> host copy of
        # the winners.  The host consumes step i-1's records (event wait) while step i is in flight - how a serving loop
        # that keeps the GPU busy reads its results; the final barrier waits for the last step.
        d_states.copy_(h_states, non_blocking=True)
        out = sp.plan_device(d_states, B, seed=2000 + i)
        h_outs[i & 1].copy_(out, non_blocking=True)
        e2e_done[i & 1].record(stream)
        if i > 0:
            e2e_done[(i - 1) & 1].synchronize()
        return h_outs[i & 1]

    def timed(step_fn, steps, per_step_events):
        evs = []
        barrier()
        t_wall = time.perf_counter()
        for i in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn(i)
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        wall = time.perf_counter() - t_wall
        ms = [a.elapsed_time(b) for a, b in evs]
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)                    # max over ranks, on the device
        return float(tot.item()), ms, wall

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        clocks.wait_ready()               # the poller has produced its first sample before the warm-up (no idle gap after it)
    barrier()
    for i in range(args.warmup):
        step_resident(i)
        step_e2e(i)
    barrier()
    # ---- resident-input throughput (value) + kernel-only duration for the roofline ----
    pl.enable_timing(True)
    launches0 = pl.kernel_launches
    kernel_ms = []
    evs_tot, ms_list, wall = 0.0, [], 0.0
    barrier()
    clocks.begin()
    t_wall = time.perf_counter()
    pairs = []
    for i in range(args.steps):        # enqueued back to back: the host does not synchronise inside the timed region
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_resident(i)
        e1.record(stream)
        pairs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - t_wall
    clocks.end()
    launches = pl.kernel_launches - launches0
    for i in range(min(args.steps, 10)):   # duration of the rollout kernel alone (the handle's own event pair), same conditions
        flush.zero_()
        step_resident(args.steps + i)
        torch.cuda.synchronize()
        kernel_ms.append(pl.last_kernel_ms())
    ms_list = [a.elapsed_time(b) for a, b in pairs]
    tot = torch.tensor([sum(ms_list)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = float(tot.item())
    pl.enable_timing(False)
    # per rank: the rollout kernel alone and the whole step - at N > 1 the step of the job is the slowest GPU's (kernel-time spread
    # between the GPUs of a box is what the scaling efficiency loses; the record exchange itself is a few microseconds)
    per_rank = None
    if world > 1:
        mine = torch.tensor([float(np.mean(kernel_ms)), sum(ms_list) / len(ms_list)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"kernel_ms": [float(t[0].item()) for t in allr], "step_ms": [float(t[1].item()) for t in allr]}
    # ---- end-to-end through the host-buffer path ----
    clocks.begin()
    e2e_ms, _, _ = timed(step_e2e, args.steps, True)
    clocks.end()
    clk = clocks.stop() if rank == 0 else None
    # ---- sustained leg: seconds of back-to-back batched plans (no L2 flush, no host synchronisation inside), clocks and
    # power sampled inside it; quoted against the sustained cuBLAS figure of MEASURED_PEAKS.json ----
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(50, int(args.sustain_s * 1.15 / (total_ms / args.steps * 1e-3)))
        sclk = ClockSampler(local)
        if rank == 0:
            sclk.start()
            sclk.wait_ready()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sclk.begin()
        e0.record(stream)
        for i in range(n_sus):
            step_resident(5000 + i)
        e1.record(stream)
        barrier()
        sclk.end()
        sus_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sus_ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            sus_s = float(sus_ms.item()) * 1e-3
            c = sclk.stop()
            p = os.path.join(ROOT, "MEASURED_PEAKS.json")
            sus_peak = float(json.load(open(p)).get("bf16_tflops_sustained", 1408.5)) if os.path.exists(p) else 1400.0
            tr_s = B * K_PATHS * HORIZON * n_sus / sus_s
            tf = tr_s / world * FLOP_PER_TRANSITION / 1e12
            sustained = {"seconds": sus_s, "steps": n_sus, "transitions_per_s": tr_s, "ms_per_step": 1e3 * sus_s / n_sus,
                         "tflops_algorithmic_per_gpu": tf, "peak_sustained_tflops": sus_peak, "frac_sustained": tf / sus_peak,
                         "frac_burst": tf / measured_peaks()[0], "sm_mhz_median": c.get("sm_mhz"), "sm_mhz_min": c.get("sm_min_mhz"),
                         "power_w_median": c.get("power_w_median"), "power_w_max": c.get("power_w_max"),
                         "sw_power_cap_share": c.get("sw_power_cap_share"), "reasons": c.get("reasons"), "clock_samples": c.get("samples"),
                         "note": "back-to-back batched plans on one stream, weights L2-resident, device time from one CUDA event pair "
                                 "around the whole run (max over ranks); the rollout kernel is > 99 % of each step"}

    sp.planner.close()
    del flush
    configs4 = None
    if not args.no_configs4:
        try:
            configs4 = configs4_leg(torch, dist, world, rank, local)
        except Exception as e:      # never breaks the contract line
            configs4 = {"error": repr(e)}
    transitions_per_step = B * K_PATHS * HORIZON                          # all ranks together
    value = transitions_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = transitions_per_step * args.steps / (e2e_ms * 1e-3)
    if rank == 0:
        peak, peak_src = measured_peaks()
        k_ms = float(np.mean(kernel_ms))
        per_gpu_transitions = transitions_per_step / world
        achieved = per_gpu_transitions * FLOP_PER_TRANSITION / (k_ms * 1e-3) / 1e12
        line = {"metric": "dynamics transitions/s (HalfCheetah random-shooting MPC)", "value": value, "unit": "transitions/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "plans_per_s": B * args.steps / (total_ms * 1e-3),
                "config": workload_config(world),
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_unit": "bytes per launch (dram read+write)",
                             "traffic_source": NCU_TRAFFIC_SOURCE,
                             "kernel": "bcmpc::tmx::rollout_tmem_kernel<2, 8, 0>", "kernel_ms": k_ms,
                             "flop_per_transition": FLOP_PER_TRANSITION, "peak_source": peak_src,
                             "note": "algorithmic FLOPs only (26-500-500-20); padded tiles (32/512/512/32) are not credited"},
                "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": B * 20 * 8,
                        "d2h_bytes_per_step": B * 80, "plans_per_s": B * args.steps / (e2e_ms * 1e-3)},
                "record_exchange": getattr(sp, "exchange", None) if world > 1 else None, "per_rank": per_rank,
                "gpu_launches": int(launches), "clocks": clk, "wall_s_timed_region": wall, "sustained": sustained, "configs4": configs4,
                "step_ms": {"first": ms_list[0], "min": min(ms_list), "median": sorted(ms_list)[len(ms_list) // 2], "max": max(ms_list)}}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        if world == 1 and not args.no_extras:
            try:
                line["extras"] = extras()
            except Exception as e:  # extras never break the contract line
                line["extras"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-configs4", action="store_true", help="skip the BASELINE.json configs[4] leg (3 x 1024 net, K = 2^20 x H = 30)")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="seconds of the sustained leg (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
