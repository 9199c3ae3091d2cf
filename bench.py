#!/usr/bin/env python
"""Headline benchmark: agent-steps/sec of the batched env rollout hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

One "step" = one pass of the rollout hot path over one batch: T lockstep env steps of n_envs
Checkers-v0 instances written straight into the rollout buffer (`mm_checkers_rollout`, one launch)
followed by the GAE/return reverse scan over that buffer (`mm_gae_returns`, one launch).  Actions are
synthetic (uniform in [0,5), seed 42) and resident in HBM before the timed region.  Every rank owns
its own env shard (no data-path collective): scaling is weak.

Workloads (per GPU):
  mappo_rollout_16384x128  BASELINE.json configs[2]: 16 384 envs x T=128, A=2, D=47 (default; the
                           configuration the agent-steps/sec metric and its roofline are quoted on)
  qmix_rollout_4096x100    BASELINE.json configs[1]: 4 096 envs x T=100 (episode written to replay slots)

Prints ONE JSON line (rank 0).  Besides the contract keys it carries: `roofline` (CUDA-event kernel times inside the timed
region against MEASURED_PEAKS.json; sub-objects `gae`, `learner`, `mappo_iteration`, `qmix_toplevel` hold the rooflines of the
other kernels), a compact `summary` as the LAST key (QMIX train samples/sec, ms per update, MAPPO iteration split ...), `e2e` (the same rollout for a caller holding HOST arrays:
BatchedCheckers.rollout_host, H2D / kernels / D2H pipelined in mm_checkers_rollout_host, wall clock), `cpu_baseline`,
`clocks` (a separate NVML sampling process), and the sections `policy_rollout`, `learner` (QMIX train samples/sec),
`mappo_iteration`, `env_switch2`, `qmix_toplevel` (skipped with --no-learner).  `--impl reference` times the CPU restatement
of the reference env (oracle/, pure Python like ma-gym itself) on all host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "mappo_rollout_16384x128": dict(n_envs=16384, T=128, full_observable=False, max_steps=100),
    "qmix_rollout_4096x100": dict(n_envs=4096, T=100, full_observable=False, max_steps=100),
}
AGENTS = 2
# algorithmic bytes per env-step (D=47): actions 8 + state 40 r + 40 w + obs 376 + rew 8 + done 2 = 474.
# The T-step kernel keeps the env state in registers (24 B per env per LAUNCH), so the bytes it must move per env-step are
# actions 8 + obs 376 + rewards 8 + dones 2 = 394: `roofline.frac` is quoted on those, the survey figure is reported beside it.
ENV_STEP_BYTES_SURVEY = 474
ENV_STEP_BYTES = 394
GAE_BYTES_PER_ELEM = 16
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # fp32 FFMA peak of one B200 at clocks.max.sm
METRIC = "agent-steps/sec batched env rollout"
UNIT = "agent-steps/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def tf32_peak_tflops():
    """Dense TF32 tensor peak = half the MEASURED sustained bf16 figure (MEASURED_PEAKS.json has no TF32 entry)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["bf16_tflops_sustained"]) / 2, "measured bf16 sustained / 2"
    return 1400.0 / 2, "fallback bf16 sustained / 2"


_SAMPLER_CHILD = r"""
import subprocess, sys, time
idx = int(sys.argv[1])
REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
try:
    import pynvml as n
    n.nvmlInit()
    h = n.nvmlDeviceGetHandleByIndex(idx)
    print("max", float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)), "nvml", flush=True)
    while True:
        mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        names = ",".join(v for b, v in REASONS.items() if mask & b) or "-"
        print(time.monotonic(), float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), names, flush=True)
        time.sleep(0.001)
except ImportError:
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    first = True
    while True:
        out = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", str(idx)],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if first:
            print("max", float(f[1]), "nvidia-smi", flush=True)
            first = False
        names = ",".join(nm for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                              "sw_power_cap"], f[2:6]) if v.lower().startswith("active")) or "-"
        print(time.monotonic(), float(f[0]), names, flush=True)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    The timed region of a short run lasts milliseconds, and a sampling thread inside this process starves
    behind the GIL while the main thread enqueues launches.  So the sampler is a separate PROCESS polling NVML
    (the library nvidia-smi itself uses) at about 1 kHz and stamping every sample with CLOCK_MONOTONIC, which
    is system-wide: the parent brackets the timed region with the same clock and keeps the samples inside it.
    Without pynvml the child falls back to an `nvidia-smi --query-gpu=...` loop.
    """

    def __init__(self, gpu_index):
        import tempfile
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = gpu_index
        if visible:
            ids = visible.split(",")
            if gpu_index < len(ids) and ids[gpu_index].strip().isdigit():
                idx = int(ids[gpu_index])
        self.t_begin = self.t_end = None
        self._out = tempfile.NamedTemporaryFile("w+", suffix=".clocks", delete=False)
        self._proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_CHILD, str(idx)], stdout=self._out,
                                      stderr=subprocess.DEVNULL)
        t0 = time.monotonic()
        while time.monotonic() - t0 < 10.0 and self._proc.poll() is None:      # wait for the first sample
            if os.path.getsize(self._out.name) > 40:
                break
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.monotonic()

    def mark_end(self):
        self.t_end = time.monotonic()

    def stop(self):
        time.sleep(0.003)
        self._proc.terminate()
        try:
            self._proc.wait(timeout=10)
        except subprocess.TimeoutExpired:
            self._proc.kill()
        with open(self._out.name) as f:
            lines = f.read().splitlines()
        os.unlink(self._out.name)
        max_mhz, source, samples, timed, reasons, timed_reasons = None, None, [], [], set(), set()
        for ln in lines:
            f = ln.split()
            if len(f) != 3:
                continue
            if f[0] == "max":
                max_mhz, source = float(f[1]), f[2]
                continue
            t, mhz = float(f[0]), float(f[1])
            names = set() if f[2] == "-" else set(f[2].split(","))
            samples.append(mhz)
            reasons |= names
            if self.t_begin is not None and self.t_end is not None and self.t_begin <= t <= self.t_end:
                timed.append(mhz)
                timed_reasons |= names
        use = timed if timed else samples
        return {"sm_mhz": statistics.median(use) if use else None, "sm_max_mhz": max_mhz,
                "reasons": sorted(timed_reasons if timed else reasons), "samples": len(samples),
                "samples_in_timed_region": len(timed), "sm_mhz_min_timed_region": min(timed) if timed else None,
                "reasons_warmup_and_timed": sorted(reasons),
                "window": "sm_mhz / reasons: samples inside the timed region (separate sampler process, ~1 kHz, "
                          "CLOCK_MONOTONIC stamps)", "source": source}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def _py_env_worker(args):
    n_envs, steps, seed, max_steps = args
    import random
    from oracle.ma_gym_ref import CheckersRef
    rnd = random.Random(seed)
    envs = [CheckersRef(full_observable=False, max_steps=max_steps) for _ in range(n_envs)]
    for e in envs:
        e.reset()
    t0 = time.perf_counter()
    for s in range(steps):
        if s % max_steps == 0 and s:
            for e in envs:
                e.reset()
        for e in envs:
            e.step([rnd.randrange(5), rnd.randrange(5)])
    return n_envs * steps, time.perf_counter() - t0


def cpu_python_port(cores, budget_s, max_steps=100):
    """Reference-style CPU rollout: the pure-Python env restatement, one process per core."""
    import multiprocessing as mp
    n, s = 8, 200
    cnt, dt = _py_env_worker((n, s, 0, max_steps))           # calibrate on one core
    rate1 = cnt / dt
    steps = max(100, int(rate1 * budget_s / n))
    if cores == 1:
        cnt, dt = _py_env_worker((n, steps, 1, max_steps))
        return cnt * AGENTS / dt, f"{n} envs x {steps} steps, random actions, 1 process"
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_py_env_worker, [(n, steps, 100 + i, max_steps) for i in range(cores)])
        wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    return total * AGENTS / wall, f"{cores} processes x {n} envs x {steps} steps, random actions"


def cpu_c_port(budget_s, n_envs=4096, max_steps=100):
    """The C restatement (oracle/magym_ref.c) on one core -- a much faster CPU than the reference has."""
    import numpy as np
    from oracle.magym_c import CheckersBatchRef
    ref = CheckersBatchRef(n_envs, False, max_steps, -0.01, auto_reset=True)
    ref.reset()
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 5, size=(16, n_envs, 2), dtype=np.int32)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        ref.step(acts[steps % 16])
        steps += 1
    return n_envs * steps * AGENTS / (time.perf_counter() - t0)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    per_step_s = min(2.0, 60.0 / max(1, args.steps))         # the whole run stays near a minute whatever K is
    for _ in range(args.warmup):
        cpu_python_port(cores, min(0.5, per_step_s))
    sample = ""
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sample = cpu_python_port(cores, per_step_s)
        vals.append(v)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": (time.perf_counter() - t0) * 1e3 / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f32",
        "data": "synthetic", "config": {"workload": args.workload, **wl, "agents": AGENTS,
                                          "note": "pure-Python env restatement (ma-gym is pure Python), "
                                                  "each step a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def policy_rollout_section(dev, steps):
    """Rollout WITH the policy in the loop: QMIX agent network (GRU, hidden 64, tcgen05) -> eps-greedy -> Checkers
    step, T = 100 steps of n_envs lockstep envs in ONE launch (mm_qmix_rollout_tc), episodes written to replay
    layout.  This is what collect_rollout (offpolicy/runner/shared/magym_runner.py:39-141) does one env-step at a time."""
    import argparse as ap
    import torch
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.nn import AgentQFunction
    from minimarl_b200.rollout import qmix_rollout
    args = ap.Namespace(hidden_size=64, use_rnn_layer=True, use_orthogonal=True, gain=0.01, layer_N=1, recurrent_N=1,
                        use_feature_normalization=True, use_ReLU=True, use_conv1d=False, fused_precision="tf32")
    net = AgentQFunction(args, 47, 5, dev)
    out = {}
    for n_envs in (4096, 9472, 16384):
        T = 100
        env = BatchedCheckers(n_envs, max_steps=T, device=dev)
        z = dict(dtype=torch.float32, device=dev)
        bufs = dict(obs=torch.zeros((T + 1, n_envs, 2, 47), **z), act_idx=torch.empty((T, n_envs, 2), dtype=torch.int32, device=dev),
                    act_onehot=torch.empty((T, n_envs, 2, 5), **z), rewards=torch.empty((T, n_envs, 2), **z),
                    dones=torch.empty((T, n_envs, 2), **z), dones_env=torch.empty((T, n_envs), **z))
        run = lambda k: qmix_rollout(net, env, T, eps=0.3, seed=42, step_offset=k * T, **bufs)
        for k in range(3):
            run(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            run(3 + k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[f"n_envs_{n_envs}"] = {"ms_per_rollout": ms, "agent_steps_per_s": n_envs * T * 2 / (ms * 1e-3),
                                   "us_per_env_step_all_envs": ms * 1e3 / T,
                                   "net_TFLOPs": n_envs * 2 * T * 64000 / (ms * 1e-3) / 1e12}
    out["note"] = ("policy in the loop: agent net forward (64 kFLOP per agent-step, TF32 tcgen05) + eps-greedy + env step "
                   "+ obs/act/reward/done writes, one launch per 100-step rollout; 128 rows (64 envs) per CTA")
    return out


def mappo_iteration_section(dev, steps, n_envs=16384, T=128):
    """BASELINE configs[2] / [4]: one full MAPPO iteration -- fused policy + env rollout of n_envs x T steps into the device
    SharedReplayBuffer, bootstrap value + GAE, and R_MAPPO.train (ppo_epoch 15, one minibatch of all data chunks of
    length 8, recurrent actor / critic of hidden size 32: the defaults of mappo/_config.py:204-336 -- lr 1e-4, value_loss_coef 0.5,
    max_grad_norm 0.5, clip 0.2, entropy 0.01, huber 10 -- except data_chunk_length 8 instead of 5, which does not divide
    T = 128: SURVEY.md 8a row a9).  Runs on every rank (env
    shard per GPU); the PPO update all-reduces its gradients over NCCL when world > 1."""
    import argparse as ap
    import torch
    import torch.distributed as dist
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.runners import MappoRunner
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    world = dist.get_world_size() if dist.is_initialized() else 1
    args = ap.Namespace(
        hidden_size=32, layer_N=1, recurrent_N=1, use_feature_normalization=True, use_ReLU=True, use_orthogonal=True,
        gain=0.01, use_conv1d=False, use_recurrent_policy=True, use_naive_recurrent_policy=False,
        use_policy_active_masks=True, use_value_active_masks=True, use_popart=False, use_valuenorm=True,
        use_huber_loss=True, use_max_grad_norm=True, use_clipped_value_loss=True, use_gae=True, ppo_epoch=15,
        train_batch_size=1, data_chunk_length=8, clip_param=0.2, entropy_coef=0.01, value_loss_coef=0.5,
        max_grad_norm=0.5, huber_delta=10.0, actor_lr=1e-4, critic_lr=1e-4, opti_eps=1e-5, weight_decay=0, gamma=0.99,
        gae_lambda=0.95, max_step=T, batch_size=n_envs, fused_precision="tf32", max_episodes=1000, use_eval=False,
        eval_interval=10, use_linear_lr_decay=False, seed=1)
    env = BatchedCheckers(n_envs, max_steps=100, auto_reset=True, device=dev)
    torch.manual_seed(args.seed)                            # every rank starts from the same parameters (mappo/main.py seeds likewise)
    r = MappoRunner({"args": args, "train_env": env, "eval_env": env, "num_agents": AGENTS, "device": dev})
    r.warmup()
    r.run_episode(0)                                        # warm-up iteration (allocations, cuBLAS heuristics)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    infos, host_train = None, 0.0
    for k in range(steps):
        ev[k][0].record(); r.rollout()
        ev[k][1].record(); r.compute()
        ev[k][2].record(); t_h = time.perf_counter(); infos = r.train(); host_train += time.perf_counter() - t_h
        ev[k][3].record()
    torch.cuda.synchronize()
    ms = [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / steps for i in range(3)]
    tot = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = float(tot.item())
    agent_steps = world * n_envs * T * AGENTS
    out = {
        "metric": "agent-steps/sec through the full MAPPO iteration (rollout + GAE + PPO update)",
        "value": agent_steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "ms_per_iteration": total_ms,
        "ms_rollout": ms[0], "ms_gae_and_bootstrap": ms[1], "ms_ppo_update": ms[2],
        "rollout_agent_steps_per_s": n_envs * T * AGENTS / (ms[0] * 1e-3),
        "ppo_update_samples_per_s": world * n_envs * T * AGENTS * args.ppo_epoch / (ms[2] * 1e-3),
        "config": {"n_envs_per_gpu": n_envs, "T": T, "agents": AGENTS, "ppo_epoch": args.ppo_epoch, "num_mini_batch": 1,
                   "data_chunk_length": 8, "hidden": 32, "fused_rollout": bool(r.fused_rollout),
                   "fused_update": "trunk2 fwd/bwd (tcgen05 kind::f16, weight gradients in the kernel) + loss kernels",
                   "precision": "fp16 operands (11-bit significand, as tf32), fp32 accumulate",
                   "grad_allreduce": "NCCL, one flattened bucket per optimiser step" if world > 1 else None},
        "train_info": {k: float(v) for k, v in infos.items()},
        "peak_mem_GiB": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
    }
    # useful math of the update: (17.7 + 17.4) kFLOP per row forward, x3 with the backward, all rows, all epochs
    flops = n_envs * T * AGENTS * (17.7e3 + 17.4e3) * 3 * args.ppo_epoch
    peak_tc, peak_src = tf32_peak_tflops()
    peak_tc, peak_src = 2 * peak_tc, peak_src.replace(" / 2", "") + " (fp16 operands)"
    # HBM side of the same kernels: per network pass the forward reads the fp16 observation rows (96 B) and writes the saved operand
    # images (208 B) and the head outputs, the backward reads both again plus the head gradients
    hbm_bytes = n_envs * T * AGENTS * (96 + 208 + 20 + 96 + 208 + 20) * 2 * args.ppo_epoch
    out["roofline"] = {"bound": "tensor", "kernel": "trunk2_fwd_kernel / trunk2_bwd_kernel", "achieved": flops / (ms[2] * 1e-3) / 1e12,
                       "peak": peak_tc, "peak_source": peak_src, "unit": "TFLOP/s", "frac": flops / (ms[2] * 1e-3) / 1e12 / peak_tc,
                       "useful_flop_per_update": flops, "hbm_GBs": hbm_bytes / (ms[2] * 1e-3) / 1e9,
                       "hbm_frac": hbm_bytes / (ms[2] * 1e-3) / 1e9 / measured_peaks()[0],
                       "what_bounds": "instruction issue of the per-row epilogues (ncu: issue slots 46 % / 35 % busy, tensor pipe 6 % / 12 %)"}
    out["ranks_in_sync"] = _ranks_in_sync(list(r.policy.actor.parameters()) + list(r.policy.critic.parameters()))
    out["launch_mode"] = "PPO epoch replayed as a CUDA graph" if r.trainer._graph_state is not None else "eager launches"
    out["ms_ppo_update_host_enqueue"] = 1e3 * host_train / steps        # time R_MAPPO.train() keeps the host (no sync inside)
    if world > 1:                                           # per rank: device time of the update, host time of the call
        mine = torch.tensor([ms[2], 1e3 * host_train / steps], device=dev, dtype=torch.float64)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        out["ms_ppo_update_per_rank"] = [round(float(e[0]), 2) for e in every]
        out["ms_ppo_update_host_enqueue_per_rank"] = [round(float(e[1]), 2) for e in every]
    del r, env
    torch.cuda.empty_cache()
    return out


def switch_rollout_section(dev, steps=10):
    """Switch2-v0 (the qmix/_config.py default env, BASELINE configs[1]) stepped by mm_switch_rollout: T lockstep steps per launch.
    Algorithmic bytes per env-step (A = 2, D = 3): actions 8 + state 12 r + 12 w + obs 24 + rewards 8 + dones 2 = 66 B."""
    import torch
    from minimarl_b200 import BatchedSwitch
    peak, _ = measured_peaks()
    out = {}
    for n, T in ((4096, 100), (262144, 128)):
        env = BatchedSwitch(n, max_steps=100, step_cost=-0.01, device=dev)
        acts = torch.randint(0, 5, (T, n, 2), dtype=torch.int32, device=dev, generator=torch.Generator(dev).manual_seed(42))
        bufs = env.new_outputs(T)
        for _ in range(3):
            env.reset()
            env.rollout(T, actions=acts, out=bufs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            env.rollout(T, actions=acts, out=bufs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = n * T * 66 / (ms * 1e-3) / 1e9
        out[f"n_envs_{n}_T_{T}"] = {"ms_per_rollout": ms, "agent_steps_per_s": n * T * 2 / (ms * 1e-3),
                                    "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peak}
    out["note"] = ("4 096 envs are 64 warps: the launch is bound by the latency of 100 dependent steps (0.35 us each), not by "
                   "HBM; the roofline fraction is meaningful at the large size")
    return out


def qmix_toplevel_section(dev, episodes=3, n_envs=4096):
    """BASELINE configs[1]: the qmix/main.py loop on 4 096 lockstep Switch2-v0 envs (the qmix/_config.py default env) with
    the chunk replay on the GPU, on the fused kernels (csrc/toplevel_qnet.cu, toplevel_mix.cu): mm_toplevel_rollout plays
    the episode (behaviour Q_Net, one epsilon draw per env, Switch2 step), mm_toplevel_qnet_seq runs the target network,
    mm_toplevel_td sums cal_td_error per chunk, the chunks of live envs enter the PER as sets, then Train_dqn.train
    (update_iter 10, recurrent Mix_Net on both sides) as fused forward / BPTT kernels.  qmix/_config.py defaults except
    buffer_limit (1 000 -> 8 192 chunks) and batch_size (32 -> 4 096 chunks), scaled with the env count."""
    import argparse as ap
    import torch
    from minimarl_b200.runners import run_qmix
    torch.backends.cuda.matmul.allow_tf32 = True
    batch = 4096
    args = ap.Namespace(use_recurrent=True, buffer_limit=8192, batch_size=batch, update_iter=10, chunk_size=10, gamma=0.99,
                        eps=1e-6, alpha=0.8, beta=0.2, use_step_weight=False, update_alpha_beta=True, max_episodes=100000,
                        grad_clip_norm=5, lr=1e-3, seed=42, max_step=100, step_cost=-0.01, max_epsilon=0.9, min_epsilon=0.05,
                        epsilon_anneal_episode=60000, update_target_interval=20, rollout_graph=True, train_graph=True)
    run_qmix(args, n_envs, 1, device=dev)                     # warm-up (allocations, cuBLAS heuristics)
    timing = {}
    history, _ = run_qmix(args, n_envs, episodes + 1, device=dev, timing=timing)
    # the first episode of a run holds the one-time graph captures: report the median of the others
    roll_s = statistics.median(timing["rollout_s_per_episode"][1:])
    train_s = statistics.median(timing["train_s_per_episode"][1:])
    steps_per_episode = timing["env_steps"] / (episodes + 1) / n_envs
    samples = args.update_iter * batch * args.chunk_size
    # useful math per agent-step (D = 3): behaviour + target Q_Net, 8 544 MACs each (fp32 FFMA kernels); per sample of the
    # update: both agents, behaviour fwd + 2x bwd + target fwd, plus the recurrent mixer (~8.8 k MACs, same four passes)
    roll_flops = steps_per_episode * n_envs * 2 * 2 * 2 * 8544
    train_flops = samples * (2 * 4 * 2 * 8544 + 4 * 2 * 8800)
    return {"metric": "qmix/main.py loop, 4 096 lockstep Switch2-v0 envs, chunk replay on the GPU",
            "fused": bool(timing.get("fused")),
            "roofline": {"bound": "fp32 ffma", "peak": FFMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "rollout_achieved": roll_flops / roll_s / 1e12, "rollout_frac": roll_flops / roll_s / 1e12 / FFMA_PEAK_TFLOPS,
                         "train_achieved": train_flops / train_s / 1e12, "train_frac": train_flops / train_s / 1e12 / FFMA_PEAK_TFLOPS},
            "rollout_agent_steps_per_s": steps_per_episode * n_envs * 2 / roll_s,
            "rollout_ms_per_env_step_all_envs": roll_s * 1e3 / steps_per_episode,
            "train_samples_per_s": samples / train_s, "train_ms_per_call": train_s * 1e3,
            "first_episode_with_graph_capture_ms": {"rollout": timing["rollout_s_per_episode"][0] * 1e3,
                                                     "train": timing["train_s_per_episode"][0] * 1e3},
            "episodes": episodes, "chunk_inserts_per_episode": n_envs * args.max_step // args.chunk_size,
            "launches": "rollout: one launch per episode (policy + eps-greedy + env), one target pass, one PER insert call; "
                        "training iteration replayed as a CUDA graph",
            "config": {"n_envs": n_envs, "env": "Switch2-v0", "buffer_limit_chunks": args.buffer_limit,
                       "batch_chunks": batch, "chunk_size": args.chunk_size, "update_iter": args.update_iter},
            "train_loss": history[-1]["train_loss"], "train_score": history[-1]["train_score"]}


def learner_section(dev, steps):
    """Second half of the BASELINE metric: QMIX train samples/sec (configs[3]: 1M-transition device replay =
    10 000 episodes x T=100, batch 4 096 episodes, GRU agents, hidden 64, PER, double-Q).  One sample = one
    transition of one env (all agents) consumed by an optimiser step, so one update = 409 600 samples."""
    import argparse as ap
    import numpy as np
    import torch
    from minimarl_b200.algorithms import QMix, QMixPolicy
    from minimarl_b200.buffers import PrioritizedRecReplayBuffer
    from minimarl_b200.envs.spaces import Box, Discrete
    T, A, D, NA, S, B = 100, 2, 47, 5, 10000, 4096
    # TF32 operands / fp32 accumulation throughout the learner: the custom kernels (tcgen05 kind::tf32) and, with this
    # flag, the remaining cuBLAS GEMMs (hypernetworks, input projections, weight gradients)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    args = ap.Namespace(
        hidden_size=64, use_rnn_layer=True, use_orthogonal=True, gain=0.01, layer_N=1, recurrent_N=1,
        use_feature_normalization=True, use_ReLU=True, use_conv1d=False, prev_act_inp=False, batch_size=B,
        mixer_hidden_dim=32, hypernet_hidden_dim=64, hypernet_layers=2, use_per=True, per_eps=1e-6, per_nu=0.9,
        use_huber_loss=False, use_popart=False, lr=5e-4, tau=0.005, opti_eps=1e-5, episode_length=T,
        use_same_share_obs=True, use_double_q=True, gamma=0.99, max_grad_norm=10.0, epsilon_start=1.0,
        epsilon_finish=0.05, epsilon_anneal_time=50000, seed=1, fused_precision="tf32")
    info = {"policy_0": {"obs_space": Box(np.zeros(D), np.ones(D)), "share_obs_space": Box(np.zeros(A * D), np.ones(A * D)),
                         "act_space": Discrete(NA)}}
    buf = PrioritizedRecReplayBuffer(0.6, info, {"policy_0": [0, 1]}, S, T, True, False, device=dev)
    pb = buf.policy_buffers["policy_0"]
    g = torch.Generator(device=dev).manual_seed(42)
    pb.obs.copy_((torch.rand(pb.obs.shape, device=dev, generator=g) * 100).round() / 100)
    pb.share_obs.copy_(pb.obs.reshape(T + 1, S, A * D))
    pb.acts.copy_(torch.nn.functional.one_hot(torch.randint(0, NA, (T, S, A), device=dev, generator=g), NA).float())
    pb.rewards.copy_(torch.randn(pb.rewards.shape, device=dev, generator=g))
    pb.dones.zero_(); pb.dones_env.zero_(); pb.dones_env[-1] = 1
    pb.filled_i, pb.current_i = S, 0
    buf.update_priorities(torch.arange(S, device=dev), torch.rand(S, device=dev, generator=g) * 0.99 + 0.01, "policy_0")
    pconf = {"cent_obs_dim": A * D, "cent_act_dim": A * NA, "obs_space": info["policy_0"]["obs_space"],
             "share_obs_space": info["policy_0"]["share_obs_space"], "act_space": info["policy_0"]["act_space"]}
    torch.manual_seed(args.seed)                            # every rank starts from the same parameters
    policy = QMixPolicy(args, {"args": args, "device": dev, "num_agents": A}, pconf)
    trainer = QMix(args, A, B, {"policy_0": policy}, lambda a: "policy_0", device=dev, episode_length=T)

    def update():
        sample = buf.sample(B, beta=0.4, p_id="policy_0")
        info_, prio, idx = trainer.train_policy_on_batch(sample)
        buf.update_priorities(idx, prio, "policy_0")
        trainer.soft_target_updates()
        return info_

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    for _ in range(3):
        update()
    ms_eager, _ = timed(update, max(3, steps // 4))          # launched from Python one kernel at a time (host-bound)
    # component timings (each in isolation, after the eager updates and before the graph's private pool exists)
    ms_sample, sample = timed(lambda: buf.sample(B, beta=0.4, p_id="policy_0"), 5)
    ms_gather, _ = timed(lambda: buf._gather(sample[8]), 10)          # the gather launch alone (indices already drawn)
    stacked = torch.cat(list(sample[0]["policy_0"]), dim=-2)
    from minimarl_b200.nn.agent_q import agentq_forward
    net = policy.q_network
    h0 = torch.zeros(A * B, 64, device=dev)
    with torch.no_grad():
        ms_agent_tc, _ = timed(lambda: agentq_forward(net.packed_tc(), stacked, h0, D, 64, NA, tensor_cores=True), 5)
        ms_agent, _ = timed(lambda: agentq_forward(net.packed(), stacked, h0, D, 64, NA), 5)
        qs = torch.randn((T, B, A), device=dev)
        ms_mixer, _ = timed(lambda: trainer.target_mixer(qs, sample[1]["policy_0"][:-1]), 5)
    graphed = None
    if True:
        # the same update as ONE CUDA graph (algorithms.qmix.GraphedUpdate): ~170 launches take the host ~6 ms to issue, the
        # device needs < 5; with several GPUs the NCCL gradient all-reduce is captured with the rest (one graph per rank)
        from minimarl_b200.algorithms.qmix import GraphedUpdate
        graphed = GraphedUpdate(trainer, buf, B)
        eager_update, update = update, (lambda: graphed(0.4))
        for _ in range(3):
            update()
    if world > 1:
        dist.barrier()
    ms_update, info_ = timed(update, steps)
    if world > 1:                       # weak scaling: every rank trains on its own batch, gradients all-reduced (NCCL)
        t_ = torch.tensor([ms_update], device=dev, dtype=torch.float64)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        ms_update = float(t_.item())
    rows = (T + 1) * A * B
    peak, _ = measured_peaks()
    gather_bytes = 2 * 81952 * B
    return {
        "metric": "QMIX train samples/sec", "value": world * B * T / (ms_update * 1e-3), "unit": "samples/s",
        "n_gpus": world, "grad_allreduce": "NCCL all-reduce of one flattened fp32 bucket per update" if world > 1 else None,
        "config": {"buffer_episodes": S, "episode_length": T, "batch_episodes_per_gpu": B, "agents": A, "hidden": 64,
                   "per": True, "double_q": True, "precision": "tf32 / fp16 operands (11-bit significand), fp32 accumulate; every kernel of the update is this repo's "
                                "(no cuBLAS / cutlass launch)"},
        "ms_per_update": ms_update, "ms_per_update_eager_launches": ms_eager,
        "launch_mode": "one CUDA graph per update (GraphedUpdate)" if graphed is not None else "eager launches",
        "ranks_in_sync": _ranks_in_sync(trainer.parameters),
        "loss": float(info_["loss"]),
        "sample_gather": {"ms": ms_sample, "algorithmic_GBps": gather_bytes / (ms_sample * 1e-3) / 1e9,
                          "frac_of_measured_hbm": gather_bytes / (ms_sample * 1e-3) / 1e9 / peak,
                          "note": "PER draw (uniforms, mm_per_sample) + mm_replay_gather; 2 x 81 952 B per sampled episode"},
        "gather_only": {"ms": ms_gather, "algorithmic_GBps": gather_bytes / (ms_gather * 1e-3) / 1e9,
                        "frac_of_measured_hbm": gather_bytes / (ms_gather * 1e-3) / 1e9 / peak,
                        "note": "mm_replay_gather for 4 096 already-drawn episode indices"},
        "agentq_forward_tcgen05": {"ms": ms_agent_tc, "rows_x_steps": rows,
                                   "TFLOPs": rows * 64000 / (ms_agent_tc * 1e-3) / 1e12,
                                   "note": "fused tcgen05 kernel (TF32 in, fp32 TMEM accumulators): target / greedy / "
                                           "rollout passes; 64 kFLOP per (row, step); 64 CTAs of 128 rows"},
        "agentq_forward_fp32": {"ms": ms_agent, "TFLOPs": rows * 64000 / (ms_agent * 1e-3) / 1e12,
                                "note": "fused FFMA kernel, the 1e-5 parity path"},
        "mixer_forward": {"ms": ms_mixer, "rows": T * B, "TFLOPs": T * B * 54528 / (ms_mixer * 1e-3) / 1e12},
        "roofline": _learner_roofline(ms_update, rows, T * B, ms_gather, gather_bytes, peak),
    }


def _ranks_in_sync(params):
    """Multi-GPU: the learners start from the same seed and average their gradients, so the parameters must stay bit-identical
    on every rank (also through the captured NCCL all-reduce of the graph replays).  None on one GPU."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return None
    flat = torch.cat([p.detach().reshape(-1).double() for p in params])
    mine = torch.stack([flat.sum(), flat.abs().sum(), (flat * flat).sum()])
    every = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(every, mine)
    return bool(all(torch.equal(e, every[0]) for e in every) and torch.isfinite(mine).all())


def _learner_roofline(ms_update, agent_rows, mixer_rows, ms_gather, gather_bytes, hbm_peak):
    """Useful math of one update: agent net 64 kFLOP per (row, step) x (live fwd + 2x bwd + target fwd), mixer
    54.5 kFLOP per (t, b) x the same four passes; against the TF32 tensor peak.  HBM side: the replay gather."""
    flops = agent_rows * 64000 * 4 + mixer_rows * 54528 * 4
    peak_tc, peak_src = tf32_peak_tflops()
    return {"bound": "tensor", "kernel": "train_policy_on_batch (all kernels of one update)", "achieved": flops / (ms_update * 1e-3) / 1e12,
            "peak": peak_tc, "peak_source": peak_src, "unit": "TFLOP/s", "frac": flops / (ms_update * 1e-3) / 1e12 / peak_tc,
            "useful_flop_per_update": flops,
            "hbm_gather": {"achieved": gather_bytes / (ms_gather * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": gather_bytes / (ms_gather * 1e-3) / 1e9 / hbm_peak}}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(gpu_index):
    """Restricts this rank to the CPU cores NVML reports as local to its GPU (intersected with the cores the process may use),
    so that the page-locked host buffers of the e2e leg are first-touched on the GPU's NUMA node.  Returns the core count, or
    None when NVML is unavailable / the intersection is empty (then nothing changes)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = gpu_index
        if visible:
            ids = visible.split(",")
            if gpu_index < len(ids) and ids[gpu_index].strip().isdigit():
                idx = int(ids[gpu_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        local = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0) & local
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return len(allowed)
    except Exception:
        return None


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import minimarl_b200
    from minimarl_b200.ops import gae_returns

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: minimarl_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)
    if world > 1:
        # stdout carries the ONE JSON line only: NCCL prints its version banner to fd 1 when NCCL_DEBUG is set, so fd 1 points
        # at stderr while the communicator is created (init + one warm-up all-reduce)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    n, T, A = wl["n_envs"], wl["T"], AGENTS
    D = 47 * (2 if wl["full_observable"] else 1)
    env = minimarl_b200.BatchedCheckers(n, full_observable=wl["full_observable"], max_steps=wl["max_steps"],
                                        step_cost=-0.01, device=dev)
    # rollout buffer laid out like SharedReplayBuffer (mappo/runner/shared/shared_buffer.py:48-77)
    obs = torch.zeros((T + 1, n, A, D), device=dev)
    rewards = torch.zeros((T, n, A, 1), device=dev)
    dones = torch.zeros((T, n, A), dtype=torch.uint8, device=dev)
    value_preds = torch.randn((T + 1, n, A, 1), device=dev, generator=torch.Generator(dev).manual_seed(42 + rank))
    masks = torch.ones((T + 1, n, A, 1), device=dev)
    returns = torch.zeros((T + 1, n, A, 1), device=dev)
    advantages = torch.zeros((T, n, A, 1), device=dev)
    actions = torch.randint(0, 5, (T, n, A), dtype=torch.int32, device=dev,
                            generator=torch.Generator(dev).manual_seed(42 + rank))

    def step():
        env.reset(out=obs[0])
        env.rollout(T, actions=actions, out=(obs[1:], rewards.view(T, n, A), dones))
        gae_returns(rewards, value_preds, masks, returns, 0.99, 0.95, use_gae=True,
                    denorm=(1.1401754, 0.1), advantages=advantages)
    launches_per_step = 3

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    if sampler:
        sampler.mark_begin()
    start.record()
    for k in range(args.steps):
        env.reset(out=obs[0])
        ev[3 * k].record()
        env.rollout(T, actions=actions, out=(obs[1:], rewards.view(T, n, A), dones))
        ev[3 * k + 1].record()
        gae_returns(rewards, value_preds, masks, returns, 0.99, 0.95, use_gae=True,
                    denorm=(1.1401754, 0.1), advantages=advantages)
        ev[3 * k + 2].record()
    end.record()
    sync_all()
    if sampler:
        sampler.mark_end()
    ms_total = start.elapsed_time(end)
    rollout_ms = statistics.mean(ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(args.steps))
    gae_ms = statistics.mean(ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(args.steps))
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())

    agent_steps_per_step = n * T * A * world
    value = agent_steps_per_step * args.steps / (ms_total * 1e-3)

    # ---- e2e: the numpy-facing env API with HOST buffers, every copy inside the timed region ----
    # one e2e step = the same rollout as above for a caller holding host arrays: reset, then rollout_host (actions
    # [T,N,A] from page-locked host memory -> obs / rewards / dones of all T steps back in host memory), H2D, kernels
    # and D2H pipelined chunk by chunk inside mm_checkers_rollout_host
    hb = env.host_buffers(T)
    host_acts = hb["act"].numpy()
    host_acts[...] = actions.cpu().numpy()
    e2e_steps = max(1, min(args.steps, 20))
    e2e_by_transport = {}
    for transport in ("direct", "compact"):              # both paths timed; the public call (transport="auto") keeps the faster
        for _ in range(3):
            env.reset()
            env.rollout_host(host_acts, transport=transport)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            env.reset()
            o_h, r_h, d_h = env.rollout_host(host_acts, transport=transport)   # returns after the last output landed
        e2e_by_transport[transport] = time.perf_counter() - t0
    for _ in range(3):                                   # the call a user makes: auto picks per host / rank count
        env.reset()
        env.rollout_host(host_acts)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        env.reset()
        o_h, r_h, d_h = env.rollout_host(host_acts)
    e2e_s = time.perf_counter() - t0
    e2e_transport = hb["transport"]
    e2e_checksum = float(r_h.sum()) + float(o_h[-1].sum())   # the host result is read
    # the one-env-step-per-call variant (step_host: H2D, kernel, D2H, sync every env step), for comparison
    sw_steps = min(T, 32)
    env.reset()
    for t in range(3):
        env.step_host(host_acts[t])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for t in range(sw_steps):
        env.step_host(host_acts[t])
    sw_s = time.perf_counter() - t1
    if world > 1:
        t = torch.tensor([e2e_s, sw_s, e2e_by_transport["direct"], e2e_by_transport["compact"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, sw_s = float(t[0].item()), float(t[1].item())
        e2e_by_transport = {"direct": float(t[2].item()), "compact": float(t[3].item())}
    e2e_value = n * A * T * e2e_steps * world / e2e_s
    e2e_stepwise = n * A * sw_steps * world / sw_s
    h2d_step, d2h_step = env.rollout_host_bytes(T, e2e_transport or "direct")

    learner = mappo_it = None
    if not args.no_learner:             # on every rank: the update all-reduces its gradients when world > 1
        try:
            learner = learner_section(dev, args.learner_steps)
        except Exception as e:          # the headline line must survive a learner failure; report it
            learner = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
        s2 = ClockSampler(local) if rank == 0 else None      # clocks / throttle reasons while the tensor-core-heavy update runs
        if s2:
            s2.mark_begin()
        try:
            mappo_it = mappo_iteration_section(dev, max(2, args.learner_steps // 2))
        except Exception as e:
            mappo_it = {"error": f"{type(e).__name__}: {e}"}
        if s2:
            s2.mark_end()
            c2 = s2.stop()
            mappo_it["clocks"] = {k: c2[k] for k in ("sm_mhz", "sm_mhz_min_timed_region", "sm_max_mhz", "reasons", "samples_in_timed_region")}

    if rank == 0:
        peak, peak_kind = measured_peaks()
        achieved = n * T * ENV_STEP_BYTES / (rollout_ms * 1e-3) / 1e9
        achieved_survey = n * T * ENV_STEP_BYTES_SURVEY / (rollout_ms * 1e-3) / 1e9
        gae_achieved = n * T * A * (GAE_BYTES_PER_ELEM + 4) / (gae_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(args.workload, {}).get("checkers_rollout_kernel")
        cores = os.cpu_count() or 1
        cpu_v, cpu_sample = cpu_python_port(1, args.cpu_seconds)
        cpu_c = cpu_c_port(min(2.0, args.cpu_seconds))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32/f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_envs_per_gpu": n, "T": T, "agents": A, "obs_dim": D,
                       "env": "Checkers-v0", "actions": "uniform int32 [0,5) seed 42, resident in HBM",
                       "step": "reset + T-step rollout into the rollout buffer + GAE/advantages",
                       "l2": f"outputs {n * T * ENV_STEP_BYTES / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)"
                       if n * T * ENV_STEP_BYTES > 126e6 else "working set below L2; buffer rewritten every step",
                       "parallelism": f"env-shard x{world}, no data-path collective",
                       "host_cores_local_to_gpu": numa},
            "roofline": {"bound": "hbm", "kernel": "checkers_rollout_kernel", "achieved": achieved, "peak": peak,
                         "peak_source": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": n * T * ENV_STEP_BYTES, "kernel_ms": rollout_ms,
                         "formula": "394 B per env-step (actions 8 + obs 376 + rewards 8 + dones 2; the env state stays in "
                                    "registers for the T steps) x n_envs x T / kernel time / measured HBM copy peak",
                         "frac_survey_474B": achieved_survey / peak,
                         "gae": {"kernel": "gae_tile_kernel", "achieved": gae_achieved, "peak": peak, "unit": "GB/s",
                                 "frac": gae_achieved / peak, "kernel_ms": gae_ms,
                                 "algorithmic_bytes_per_launch": n * T * A * (GAE_BYTES_PER_ELEM + 4)}},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port", "sample": cpu_sample,
                             "host_cores": cores, "c_port_1core": cpu_c},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                    "steps": e2e_steps, "ms_per_step": e2e_s * 1e3 / e2e_steps,
                    "pcie_GBps": (h2d_step + d2h_step) * e2e_steps / e2e_s / 1e9, "result_checksum": e2e_checksum,
                    "host_bytes_written_per_step": n * T * (A * D * 4 + A * 4 + A),
                    "transport": e2e_transport,
                    "by_transport": {k: n * A * T * e2e_steps * world / v for k, v in e2e_by_transport.items()},
                    "api": "BatchedCheckers.reset(); BatchedCheckers.rollout_host(numpy actions [T,N,A]) -> numpy obs / rewards / "
                           "dones of all T steps in page-locked host arrays; transport auto: direct = float32 observations over "
                           "PCIe (mm_checkers_rollout_host), compact = one 64-bit word per agent-step over PCIe, expanded "
                           "bit-identically by host threads (mm_checkers_rollout_host_compact); wall clock around the calls",
                    "stepwise_value": e2e_stepwise,
                    "stepwise_api": "BatchedCheckers.step_host: one env step per call, synchronised every step"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if not args.no_learner:
            sections = {}
            for name, fn in (("policy_rollout", lambda: policy_rollout_section(dev, args.learner_steps)),
                             ("env_switch2", lambda: switch_rollout_section(dev)),
                             ("qmix_toplevel", lambda: qmix_toplevel_section(dev))):
                try:
                    sections[name] = fn()
                except Exception as e:
                    sections[name] = {"error": f"{type(e).__name__}: {e}"}
            line["policy_rollout"] = sections["policy_rollout"]
            line["env_switch2"] = sections["env_switch2"]
            line["learner"] = learner
            line["mappo_iteration"] = mappo_it
            line["qmix_toplevel"] = sections["qmix_toplevel"]
            # The second half of the BASELINE metric and the other configs, compact, where the driver's record keeps them:
            # inside `roofline` (kept whole) and as the LAST key of the line (the driver's stdout tail).
            g = lambda d, *ks: (d or {}).get(ks[0]) if len(ks) == 1 else g((d or {}).get(ks[0]), *ks[1:])
            summary = {
                "qmix_train_samples_per_s": g(learner, "value"), "qmix_ms_per_update": g(learner, "ms_per_update"),
                "qmix_update_tflops_frac_of_tf32_peak": g(learner, "roofline", "frac"),
                "qmix_gather_frac_of_hbm": g(learner, "roofline", "hbm_gather", "frac"),
                "mappo_iteration_ms": g(mappo_it, "ms_per_iteration"), "mappo_ms_rollout": g(mappo_it, "ms_rollout"),
                "mappo_ms_ppo_update": g(mappo_it, "ms_ppo_update"), "mappo_agent_steps_per_s": g(mappo_it, "value"),
                "mappo_update_tflops_frac_of_fp16_peak": g(mappo_it, "roofline", "frac"),
                "mappo_update_frac_of_hbm": g(mappo_it, "roofline", "hbm_frac"),
                "toplevel_qmix_rollout_agent_steps_per_s": g(sections["qmix_toplevel"], "rollout_agent_steps_per_s"),
                "toplevel_qmix_train_samples_per_s": g(sections["qmix_toplevel"], "train_samples_per_s"),
                "toplevel_qmix_train_ms_per_call": g(sections["qmix_toplevel"], "train_ms_per_call"),
                "policy_rollout_4096_agent_steps_per_s": g(sections["policy_rollout"], "n_envs_4096", "agent_steps_per_s"),
                "env_frac_of_hbm_394B": achieved / peak, "gae_frac_of_hbm": gae_achieved / peak,
                "e2e_agent_steps_per_s": e2e_value,
            }
            line["roofline"]["learner"] = g(learner, "roofline")
            line["roofline"]["mappo_iteration"] = g(mappo_it, "roofline")
            line["roofline"]["qmix_toplevel"] = g(sections["qmix_toplevel"], "roofline")
            line["e2e"]["learner_ms_per_update"] = g(learner, "ms_per_update")
            line["e2e"]["learner_samples_per_s"] = g(learner, "value")
            line["summary"] = summary
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mappo_rollout_16384x128", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU time spent on the cpu_baseline sample")
    ap.add_argument("--no-learner", action="store_true", help="skip the QMIX learner section")
    ap.add_argument("--learner-steps", type=int, default=5)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
