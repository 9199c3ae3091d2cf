"""Shared-policy on-policy runner with the reference's surface (mappo/runner/shared/base_runner.py:20-175,
magym_runner.py:17-241) for N truly independent lockstep envs.

The reference steps ONE env object ``batch_size`` times per time step (quirk Q1 of SURVEY.md Appendix B); here
``train_env`` is a ``BatchedCheckers`` with ``n_envs == args.batch_size`` boards (auto-reset on done), the actor
and critic run as one fused kernel each per step, the env kernel writes ``obs[t+1]`` / ``rewards[t]`` of the
device ``SharedReplayBuffer`` in place, returns come from the GAE kernel, and ``R_MAPPO.train`` does the update.

With the default network (hidden 32, local-observation critic, partial observations) the whole T-step rollout --
collect + env.step + insert of every step -- is ONE launch of ``mm_mappo_rollout_tc`` (``args.fused_rollout``,
default on; ``collect`` / ``insert`` stay available and are what the per-step path runs).
"""
import torch

from ..algorithms.mappo import R_MAPPO, R_MAPPOPolicy
from ..buffers import SharedReplayBuffer
from ..rollout import mappo_rollout, mappo_rollout_fusable, pack_mappo_rollout


class MAGYM_Runner:
    def __init__(self, config):
        self.args = args = config["args"]
        self.train_env, self.eval_env = config["train_env"], config["eval_env"]
        self.num_agents, self.device = config["num_agents"], config["device"]
        self.batch_size, self.episode_length = args.batch_size, args.max_step
        assert self.train_env.n_envs == self.batch_size, "one lockstep env per rollout row"
        self.hidden_size, self.recurrent_N = args.hidden_size, args.recurrent_N
        self.max_episodes, self.use_eval, self.eval_interval = args.max_episodes, args.use_eval, args.eval_interval
        self.use_linear_lr_decay = args.use_linear_lr_decay
        obs_space = self.train_env.observation_space[0]
        share_space = obs_space          # IPPO-style local critic unless use_centralized_V (mappo/_config.py:139-150)
        self.policy = R_MAPPOPolicy(args, obs_space, share_space, self.train_env.action_space[0], device=self.device)
        self.trainer = R_MAPPO(args, self.policy, device=self.device)
        self.buffer = SharedReplayBuffer(args, self.num_agents, obs_space, share_space, self.train_env.action_space[0],
                                         device=self.device)
        self._dones = torch.zeros((self.batch_size, self.num_agents), dtype=torch.uint8, device=self.device)
        self.train_infos = []
        self.fused_rollout = bool(getattr(args, "fused_rollout", True)) and mappo_rollout_fusable(self.policy, self.train_env)
        self._packed, self._packed_key = None, None
        self._seed, self._steps_done = int(getattr(args, "seed", 0)), 0

    def _rollout_params(self):
        ps = list(self.policy.actor.parameters()) + list(self.policy.critic.parameters())
        key = tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps)
        if key != self._packed_key:
            self._packed, self._packed_key = pack_mappo_rollout(self.policy.actor, self.policy.critic, out=self._packed), key
        return self._packed

    @torch.no_grad()
    def rollout(self):
        """All ``episode_length`` lockstep steps of collect / env.step / insert."""
        b, N, A = self.buffer, self.batch_size, self.num_agents
        if self.fused_rollout:
            mappo_rollout(self._rollout_params(), self.train_env, b, self.episode_length, seed=self._seed,
                          step_offset=self._steps_done)
        else:
            for step in range(self.episode_length):
                values, actions, log_probs, rnn_a, rnn_c = self.collect(step)
                self.train_env.step_into(actions.reshape(N, A).to(torch.int32).contiguous(), b.obs[step + 1],
                                         b.rewards[step].view(N, A), self._dones)
                self.insert((b.obs[step + 1], b.obs[step + 1], b.rewards[step], self._dones, values, actions, log_probs,
                             rnn_a, rnn_c))
        self._steps_done += self.episode_length

    def warmup(self):
        self.train_env.reset(out=self.buffer.obs[0])
        self.buffer.share_obs[0].copy_(self.buffer.obs[0])

    @torch.no_grad()
    def collect(self, step):
        self.trainer.prep_rollout()
        b, N, A = self.buffer, self.batch_size, self.num_agents
        flat = lambda x: x.reshape(N * A, *x.shape[2:])
        value, action, action_log_prob, rnn_state, rnn_state_critic = self.policy.get_actions(
            obs=flat(b.obs[step]), cent_obs=flat(b.share_obs[step]), rnn_states_actor=flat(b.rnn_states[step]),
            rnn_states_critic=flat(b.rnn_states_critic[step]), masks=flat(b.masks[step]))
        un = lambda x: x.reshape(N, A, *x.shape[1:])
        return un(value), un(action), un(action_log_prob), un(rnn_state), un(rnn_state_critic)

    def insert(self, data):
        next_obs, next_share_obs, rewards, dones, values, actions, action_log_probs, rnn_states, rnn_states_critic = data
        dones_env = dones.bool().all(dim=1)
        keep = (~dones_env).to(torch.float32)
        rnn_states = rnn_states * keep.view(-1, 1, 1, 1)
        rnn_states_critic = rnn_states_critic * keep.view(-1, 1, 1, 1)
        masks = keep.view(-1, 1, 1).expand(-1, self.num_agents, 1)
        self.buffer.insert(share_obs=next_share_obs, obs=next_obs, rnn_states_actor=rnn_states,
                           rnn_states_critic=rnn_states_critic, actions=actions.to(torch.float32),
                           action_log_probs=action_log_probs, value_preds=values, rewards=rewards, masks=masks)

    @torch.no_grad()
    def compute(self):
        self.trainer.prep_rollout()
        b, N, A = self.buffer, self.batch_size, self.num_agents
        flat = lambda x: x.reshape(N * A, *x.shape[2:])
        next_values = self.policy.get_values(flat(b.share_obs[-1]), flat(b.rnn_states_critic[-1]), flat(b.masks[-1]))
        b.compute_returns(next_values.reshape(N, A, 1), self.trainer.value_normalizer)

    def train(self):
        self.trainer.prep_training()
        train_infos = self.trainer.train(self.buffer)
        self.buffer.after_update()
        return train_infos

    def run_episode(self, episode):
        """One rollout of T lockstep steps + one PPO update (the body of the reference's ``run`` loop)."""
        if self.use_linear_lr_decay:
            self.policy.lr_decay(episode=episode, episodes=self.max_episodes)
        self.rollout()
        self.compute()
        infos = self.train()
        self.train_infos.append(infos)
        return infos

    def run(self, episodes=None, sink=None):
        """``sink``: a ``MetricSink``; gets ``train_infos`` (+ ``Test_Rewards``) per iteration like the reference's
        ``wandb.log(train_infos)`` (mappo/runner/shared/magym_runner.py:100-105), device tensors copied in one go."""
        self.warmup()
        for episode in range(self.max_episodes if episodes is None else episodes):
            infos = self.run_episode(episode)
            if self.use_eval and episode % self.eval_interval == 0:
                infos["Test_Rewards"] = self.eval()
            if sink is not None:
                sink.log(infos, step=episode)
        return self.train_infos

    @torch.no_grad()
    def eval(self):
        """Greedy episodes on ``eval_env`` (all its lockstep boards at once); mean team return."""
        env = self.eval_env
        N, A = env.n_envs, self.num_agents
        obs = env.reset()
        rnn = torch.zeros((N * A, self.recurrent_N, self.hidden_size), device=self.device)
        masks = torch.ones((N * A, 1), device=self.device)
        total = torch.zeros(N, device=self.device)
        alive = torch.ones(N, dtype=torch.bool, device=self.device)
        self.trainer.prep_rollout()
        for _ in range(env.max_steps):
            actions, rnn = self.policy.act(obs.reshape(N * A, -1), rnn, masks, deterministic=True)
            obs, rew, done = env.step(actions.reshape(N, A))
            total += rew.sum(1) * alive
            alive &= ~done.all(1)
            if not bool(alive.any()):
                break
        return total.mean()
