"""The training loops of ``vdn/main.py`` (:80-198) and ``qmix/main.py`` (:96-260) for N lockstep envs.

Per step: ``behavior_network.sample_action`` (one epsilon draw per env) -> lockstep env kernel ->
``target_network`` forward -> ``cal_td_error`` for every env -> every ``chunk_size`` steps the chunks of ALL
envs go into the device chunk-PER in one ``collect_samples_batch`` call (env order = the reference's add order;
``mm_heaptree_add_batch`` replays the adds, evictions included, in one launch); after each batch of episodes
``Target_Dqn.train`` (vdn) or ``Train_dqn.train`` with the recurrent ``Mix_Net`` (qmix) and the periodic hard
target copy, with the reference's epsilon / alpha / beta schedules.

``run_vdn``: Checkers-v0 with full observations (D = 94), the vdn/_config.py default.
``run_qmix``: Switch2-v0 (D = 3), the qmix/_config.py default (qmix/_config.py:13-18).

Default path (``args.fused``, on unless the shapes are outside the kernels): one ``mm_toplevel_rollout`` launch plays a whole
episode of all envs (behaviour Q_Net, epsilon-greedy, env step), one ``mm_toplevel_qnet_seq`` launch runs the target network
over the recorded next states, ``mm_toplevel_td`` sums ``cal_td_error`` per chunk, and the chunks of every env that was still
alive at the chunk's first step go into the PER as one set (``mm_heaptree_add_set`` + ``mm_rows_scatter``).  Like the reference
the run starts with an epsilon = 1 warm-up until the buffer holds ``buffer_limit`` chunks (vdn/main.py:80-125).  Differences
from the one-env reference loop that lockstep batching forces, all documented in DESIGN.md: envs that finish early idle until
the slowest env of the batch is done (their later chunks are not stored; a stored chunk may end in padding steps with
done = 1, reward 0), and a chunk never straddles two episodes.
"""
import time

import torch

from ..algorithms import toplevel_fused as tf
from ..algorithms.toplevel import Mix_Net, Prioritized_Experience_Replay, Q_Net, Target_Dqn, Train_dqn, cal_td_error
from ..envs.batched import BatchedCheckers, BatchedSwitch
from ..optim import make_adam
from ..utils.sinks import MetricSink, test_line, vdn_line


class _ChunkGraph:
    """One chunk (``chunk_size`` env steps of all lockstep envs) of the rollout body as ONE CUDA graph.

    The body -- ``sample_action``, the env kernel, the target forward, ``cal_td_error``, the chunk / score bookkeeping -- is
    ~60 small launches per step on a few thousand rows: launch-bound from Python (1.5 ms per step at 4 096 envs).  Captured
    once into static tensors and replayed per chunk it runs at GPU speed.  epsilon is a device scalar so that one graph
    serves every episode; the random draws come from the default generator, which torch registers with the capture."""

    def __init__(self, beh, tgt, env, args):
        dev, n, L = env.device, env.n_envs, args.chunk_size
        self.beh, self.tgt, self.env, self.gamma, self.L = beh, tgt, env, args.gamma, L
        z = dict(dtype=torch.float32, device=dev)
        A, D = env.n_agents, env.obs_dim
        self.state = torch.zeros((n, A, D), **z)
        self.hidden, self.target_hidden = beh.init_hidden(n), tgt.init_hidden(n)
        self.eps = torch.zeros((), **z)
        self.score = torch.zeros(n, **z)
        self.alive = torch.ones(n, dtype=torch.bool, device=dev)
        self.td_sum = torch.zeros(n, **z)
        self.chunk = [torch.zeros((n, L, A, D), **z), torch.zeros((n, L, A), **z), torch.zeros((n, L, A), **z),
                      torch.zeros((n, L, A, D), **z), torch.zeros((n, L), **z)]
        self.graph = None

    def _body(self):
        s, a_, r_, n_, d_ = self.chunk
        self.td_sum.zero_()
        state, hidden, target_hidden = self.state, self.hidden, self.target_hidden
        for l in range(self.L):
            action, next_hidden, behavior_q = self.beh.sample_action(state, hidden, self.eps)
            next_state, reward, done = self.env.step(action)
            target_q, target_hidden = self.tgt(next_state, target_hidden)
            done_all = done.all(1)
            done_env = done_all.float()
            self.td_sum += cal_td_error(action, reward, done_env, behavior_q, target_q, self.gamma)
            s[:, l], a_[:, l], r_[:, l], n_[:, l], d_[:, l] = state, action, reward, next_state, done_env
            self.score += reward.sum(1) * self.alive
            self.alive &= ~done_all
            state, hidden = next_state, next_hidden
        self.state.copy_(state)
        self.hidden.copy_(hidden)
        self.target_hidden.copy_(target_hidden)

    def begin_episode(self, obs, epsilon):
        self.state.copy_(obs)
        self.hidden.zero_(); self.target_hidden.zero_(); self.score.zero_(); self.alive.fill_(True)
        self.eps.fill_(float(epsilon))

    def capture(self):
        if self.graph is None:
            snap = [t.clone() for t in (self.state, self.hidden, self.target_hidden, self.score, self.alive, self.env.state)]
            side = torch.cuda.Stream(device=self.state.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():                 # warm-up outside the capture (allocator, cuBLAS)
                self._body()
            torch.cuda.current_stream().wait_stream(side)
            for t, v in zip((self.state, self.hidden, self.target_hidden, self.score, self.alive, self.env.state), snap):
                t.copy_(v)
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self._body()
            for t, v in zip((self.state, self.hidden, self.target_hidden, self.score, self.alive, self.env.state), snap):
                t.copy_(v)                                                 # capture does not execute, but keep it explicit
        return self

    def run_chunk(self):
        self.capture()
        self.graph.replay()
        return self.chunk, self.td_sum


def _fused_episode(beh, tgt, env, per, args, epsilon, seed, step_offset, bufs):
    """One episode of every env on the fused kernels; returns (score [N], env steps played)."""
    L, T = args.chunk_size, (args.max_step // args.chunk_size) * args.chunk_size
    out = tf.rollout_episode(beh, env, T, L, epsilon, seed=seed, step_offset=step_offset, out=bufs)
    Cn, N = T // L, env.n_envs
    q_tgt = tf.qnet_seq(tgt, out["next_states"], N, L, C_=Cn)                 # target state carried through the episode
    td_sum, valid = tf.td_chunks(q_tgt, out["rewards"], out["dones"], out["behv"], out["alive"], args.gamma)
    # chunk c of every env goes in as one set, c = 0..C-1 (the reference's add order), all from one library call
    per.collect_episode([out["states"], out["actions"], out["rewards"], out["next_states"], out["dones"]], td_sum, valid)
    return out, T * N


def _adam(params, lr, fused, capturable):
    """vdn/main.py:97 / qmix/main.py: ``optim.Adam(params, lr)``; on the fused path the flat two-launch optimiser (optim.py)."""
    if fused:
        return make_adam(params, lr=lr, eps=1e-8)
    return torch.optim.Adam(params, lr=lr, capturable=capturable)


def _run(args, env, episodes, mixer, log, timing, sink=None, test=None):
    dev, n_envs = env.device, env.n_envs
    beh = Q_Net(env.observation_space, env.action_space, args).to(dev)
    tgt = Q_Net(env.observation_space, env.action_space, args).to(dev)
    tgt.load_state_dict(beh.state_dict())
    fused = bool(getattr(args, "fused", True)) and tf.qnet_supported(beh) and env.n_agents == 2 and not (
        isinstance(env, BatchedSwitch) and env.full_observable) and args.max_step >= args.chunk_size
    if fused and not hasattr(args, "fast_tree"):
        args.fast_tree = True                                   # set-semantics PER maintenance (heaptree_batch.cu)
    per = Prioritized_Experience_Replay(args, n_agents=env.n_agents, obs_dim=env.obs_dim, device=dev)
    use_graph = bool(getattr(args, "train_graph", False))            # capture one training iteration as a CUDA graph
    if mixer:
        mix, tmix = Mix_Net(env.observation_space, args).to(dev), Mix_Net(env.observation_space, args).to(dev)
        tmix.load_state_dict(mix.state_dict())
        opt = _adam(list(beh.parameters()) + list(mix.parameters()), args.lr, fused, use_graph)
        trainer = Train_dqn(args, dev)
        train = lambda eps: trainer.train(per, beh, mix, tgt, tmix, opt, eps)
    else:
        mix = tmix = None
        opt = _adam(list(beh.parameters()), args.lr, fused, use_graph)
        trainer = Target_Dqn(per, beh, tgt, args, dev)
        train = lambda eps: trainer.train(target_network=tgt, optimizer=opt, epsilon=eps)
    L = args.chunk_size
    graph = None
    if not fused and getattr(args, "rollout_graph", True):
        graph = _ChunkGraph(beh, tgt, env, args).capture()                     # one-time capture
    history = []
    t_roll, t_train = [], []
    env_steps = 0
    seed = int(getattr(args, "seed", 0))
    bufs = None
    played = 0                                                                 # episodes played so far: the Philox counter base
    # ---- warm-up: epsilon = 1 until the buffer holds buffer_limit chunks (vdn/main.py:80-125) ---------------------------------
    if fused and getattr(args, "warm_up", True):
        while per.sum_tree.sync_n_data() < args.buffer_limit:
            bufs, n = _fused_episode(beh, tgt, env, per, args, 1.0, seed, played * args.max_step, bufs)
            played += 1
            if timing is not None:
                timing["warm_up_env_steps"] = timing.get("warm_up_env_steps", 0) + n
    for episode in range(episodes):
        epsilon = max(args.min_epsilon, args.max_epsilon - (args.max_epsilon - args.min_epsilon) *
                      (episode / args.epsilon_anneal_episode))
        if timing is not None:
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
        if fused:
            bufs, n = _fused_episode(beh, tgt, env, per, args, epsilon, seed, played * args.max_step, bufs)
            played += 1
            env_steps += n
            score = bufs["score"]
        elif graph is not None:
            graph.begin_episode(env.reset(), epsilon)
            for _ in range(args.max_step // L):
                chunk_tensors, td_sum = graph.run_chunk()
                per.collect_samples_batch(chunk_tensors, td_sum, warm_up=False)    # one chunk per env, in env order
                env_steps += n_envs * L
                if not bool(graph.alive.any()):                              # the reference stops at `all(done)`
                    break
            score = graph.score
        else:
            state = env.reset()
            hidden, target_hidden = beh.init_hidden(n_envs), tgt.init_hidden(n_envs)
            score = torch.zeros(n_envs, device=dev)
            alive = torch.ones(n_envs, dtype=torch.bool, device=dev)
            chunk = {k: [] for k in "sarnd"}
            td_sum = torch.zeros(n_envs, device=dev)
            with torch.no_grad():
                for step in range(args.max_step):
                    action, next_hidden, behavior_q = beh.sample_action(state, hidden, epsilon)
                    next_state, reward, done = env.step(action)
                    target_q, target_hidden = tgt(next_state, target_hidden)
                    done_env = done.all(1).float()
                    td_sum += cal_td_error(action, reward, done_env, behavior_q, target_q, args.gamma)
                    for k, v in zip("sarnd", (state, action, reward, next_state, done_env)):
                        chunk[k].append(v)
                    score += reward.sum(1) * alive
                    alive &= ~done.all(1)
                    env_steps += n_envs
                    if (step + 1) % L == 0:
                        stack = [torch.stack(chunk[k], 1) for k in "sarnd"]
                        per.collect_samples_batch(stack, td_sum, warm_up=False)    # one chunk per env, in env order
                        chunk = {k: [] for k in "sarnd"}
                        td_sum = torch.zeros(n_envs, device=dev)
                        if not bool(alive.any()):                                # the reference stops at `all(done)`
                            break
                    state, hidden = next_state, next_hidden
        if timing is not None:
            torch.cuda.synchronize(dev)
            t1 = time.perf_counter()
            t_roll.append(t1 - t0)
        n_data = per.sum_tree.sync_n_data() if fused else per.sum_tree.n_data
        loss = train(epsilon) if n_data >= args.batch_size else torch.tensor(float("nan"))
        if episode % args.update_target_interval == 0:
            tgt.load_state_dict(beh.state_dict())
            if mixer:
                tmix.load_state_dict(mix.state_dict())
        if timing is not None:
            torch.cuda.synchronize(dev)
            t_train.append(time.perf_counter() - t1)
        # device values stay tensors; the sink (or the history below) brings them over in one copy
        rec = {"episode": episode, "train_score": score.mean(), "train_loss": loss, "eps": epsilon,
               "alpha": per.alpha, "beta": per.beta}
        if sink is not None:
            sink.log(rec, step=episode, line=vdn_line(args.max_episodes))
        rec = MetricSink.to_host(rec)
        history.append(rec)
        if log:
            log(rec)
        if test is not None and (episode + 1) % args.test_interval == 0:           # vdn/main.py:185-190
            t_score, t_loss = test.execute(beh, tgt)
            history[-1].update({"test-score": t_score, "test-loss": t_loss})
            if sink is not None:
                sink.log({"test-score": t_score, "test-loss": t_loss}, step=episode, line=test_line, force=True)
    if timing is not None:
        timing.update(rollout_s=sum(t_roll), train_s=sum(t_train), rollout_s_per_episode=t_roll, train_s_per_episode=t_train,
                      env_steps=env_steps, chunks_in_buffer=per.sum_tree.n_data, fused=fused)
    return history, beh


def _test_module(args, env_factory, dev):
    """``Test`` (vdn/_test.py) on its own lockstep boards when the config asks for periodic evaluation."""
    if not getattr(args, "test_interval", 0) or not getattr(args, "test_episodes", 0):
        return None
    from ..algorithms.toplevel import Test
    return Test(env_factory(int(args.test_episodes)), args, dev)


def run_vdn(args, n_envs, episodes, device="cuda", log=None, timing=None, sink=None):
    """vdn/main.py on ``n_envs`` lockstep boards.  ``sink``: a ``MetricSink`` (logging / wandb lines of vdn/main.py:180-196)."""
    dev = torch.device(device)
    make = lambda n: BatchedCheckers(n, full_observable=True, max_steps=args.max_step, step_cost=args.step_cost, device=dev)
    return _run(args, make(n_envs), episodes, False, log, timing, sink, _test_module(args, make, dev))


def run_qmix(args, n_envs, episodes, device="cuda", log=None, timing=None, env_id="Switch2-v0", sink=None):
    dev = torch.device(device)
    if env_id.startswith("Switch"):
        make = lambda n: BatchedSwitch(n, n_agents=int(env_id[6]), max_steps=args.max_step, step_cost=args.step_cost, device=dev)
    else:
        make = lambda n: BatchedCheckers(n, full_observable=False, max_steps=args.max_step, step_cost=args.step_cost, device=dev)
    return _run(args, make(n_envs), episodes, True, log, timing, sink, None)
