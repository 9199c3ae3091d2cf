"""Off-policy runner with the reference's surface (offpolicy/runner/shared/base_runner.py:24-298,
magym_runner.py:9-166) for N lockstep envs.

``MAGYM_Runner(config)`` takes the same config dict (``args``, ``train_env``, ``eval_env``, ``num_agents``,
``device``); ``train_env`` is a ``BatchedCheckers`` with ``n_envs`` boards instead of one gym env.  ``run()``
= collect one batch of N episodes with the fused policy + env kernel (written straight into the device PER
buffer), then ``batch_train_q`` (sample -> ``QMix.train_policy_on_batch`` -> ``update_priorities`` -> soft /
hard target update), with the reference's counters and schedules.
"""
import numpy as np
import torch

from ..algorithms.qmix import DecayThenFlatSchedule, QMix, QMixPolicy
from ..buffers import PrioritizedRecReplayBuffer, RecReplayBuffer
from ..envs.spaces import Box
from ..rollout import EpisodeCollector


class MAGYM_Runner:
    def __init__(self, config):
        self.args = args = config["args"]
        self.train_env, self.eval_env = config["train_env"], config["eval_env"]
        self.num_agents, self.device = config["num_agents"], config["device"]
        self.algorithm_name = args.algorithm_name
        if self.algorithm_name not in ("qmix", "vdn"):
            raise NotImplementedError("the hot path covers the recurrent qmix / vdn algorithms")
        self.use_per, self.use_eval = args.use_per, args.use_eval
        self.use_soft_update = args.use_soft_update
        self.num_envs = self.train_env.n_envs
        self.batch_size, self.hidden_size, self.buffer_size = args.batch_size, args.hidden_size, args.buffer_size
        self.episode_length = args.episode_length
        self.num_env_steps = args.num_env_steps
        self.train_interval_episode = args.train_interval_episode
        self.hard_update_interval_episode = args.hard_update_interval_episode
        self.log_interval, self.eval_interval = args.log_interval, args.eval_interval
        low = np.tile(self.train_env._obs_low, self.num_agents)
        high = np.tile(self.train_env._obs_high, self.num_agents)
        self.policy_info = {"policy_0": {"cent_obs_dim": low.shape[0], "cent_act_dim": 5 * self.num_agents,
                                         "obs_space": self.train_env.observation_space[0],
                                         "share_obs_space": Box(low, high), "act_space": self.train_env.action_space[0]}}
        self.policy_ids = ["policy_0"]
        self.policy_agents = {"policy_0": list(range(self.num_agents))}
        self.total_env_steps = self.num_episodes_collected = self.total_train_steps = 0
        self.last_train_episode = self.last_eval_T = self.last_log_T = self.last_hard_update_episode = 0
        self.policies = {"policy_0": QMixPolicy(args, config, self.policy_info["policy_0"])}
        self.trainer = QMix(args, self.num_agents, self.batch_size, self.policies, lambda a: "policy_0",
                            device=self.device, episode_length=self.episode_length, vdn=self.algorithm_name == "vdn")
        num_train_episodes = (self.num_env_steps / self.episode_length) / self.train_interval_episode
        self.beta_anneal = DecayThenFlatSchedule(args.per_beta_start, 1.0, num_train_episodes, decay="linear")
        buf_args = (self.policy_info, self.policy_agents, self.buffer_size, self.episode_length, args.use_same_share_obs,
                    False, args.use_reward_normalization)
        # The reference seeds leaves 0..n-1 instead of the slots just written (rec_buffer.py:265-268, quirk Q12).  With one env
        # that only ever touches leaf 0; with N lockstep envs it would leave every slot >= N unsampled, so the runner seeds the
        # written slots.  ``faithful_per_insert=True`` restores the reference's behaviour.
        self.buffer = PrioritizedRecReplayBuffer(args.per_alpha, *buf_args, device=self.device,
                                                 fix_insert_priority=not getattr(args, "faithful_per_insert", False)) \
            if self.use_per else RecReplayBuffer(*buf_args, device=self.device)
        import torch.distributed as dist
        if self.use_per and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.buffer.sharded = True                       # env / replay shard per GPU: global PER normaliser
        self._collector = EpisodeCollector(self.train_env, self.policies["policy_0"], self.buffer, self.episode_length,
                                           seed=getattr(args, "seed", 0))
        self.train_infos, self.env_infos = [], {}
        num_warmup_episodes = max((self.batch_size, args.num_random_episodes))
        self.warmup(num_warmup_episodes)

    # -- reference surface -------------------------------------------------------------------------------
    def collect_rollout(self, explore=True, training_episode=True, warmup=False):
        self._collector.total_env_steps = self.total_env_steps
        info = self._collector.collect_rollout(explore, training_episode, warmup)
        self.total_env_steps = self._collector.total_env_steps
        if explore:
            self.num_episodes_collected += self.num_envs
        return info

    def run(self):
        self.trainer.prep_rollout()
        self.env_infos = self.collect_rollout(explore=True, training_episode=True, warmup=False)
        if ((self.num_episodes_collected - self.last_train_episode) / self.train_interval_episode) >= 1 \
                or self.last_train_episode == 0:
            self.batch_train_q()
            self.total_train_steps += 1
            self.last_train_episode = self.num_episodes_collected
        if ((self.total_env_steps - self.last_log_T) / self.log_interval) >= 1 or self.total_env_steps >= self.num_env_steps:
            self.log()
            self.last_log_T = self.total_env_steps
        if self.use_eval and ((self.total_env_steps - self.last_eval_T) / self.eval_interval) >= 1:
            self.eval()
            self.last_eval_T = self.total_env_steps
        return self.total_env_steps

    def warmup(self, num_warmup_episodes):
        self.trainer.prep_rollout()
        rewards = []
        for _ in range((num_warmup_episodes // self.num_envs) + 1):
            rewards.append(self.collect_rollout(explore=True, training_episode=False, warmup=True)["average_episode_rewards"])
        self.warmup_reward = torch.stack(rewards).mean()

    def _graphed_update(self):
        """The whole update (PER draw, gather, train_policy_on_batch, priority write-back, soft target update) as one CUDA graph
        replay (algorithms.qmix.GraphedUpdate) once the buffer is full -- its length is part of the captured sampling kernel --
        for the configuration the reference defaults to: one policy, PER, soft target updates."""
        if not (getattr(self.args, "train_graph", True) and self.use_per and self.use_soft_update and len(self.policy_ids) == 1
                and torch.device(self.device).type == "cuda" and hasattr(self.trainer, "train_policy_on_batch")
                and len(self.buffer) == self.buffer.policy_buffers[self.policy_ids[0]].buffer_size):
            return None
        g = getattr(self, "_graphed", None)
        if g is None:
            from ..algorithms.qmix import GraphedUpdate
            g = self._graphed = GraphedUpdate(self.trainer, self.buffer, self.batch_size, self.policy_ids[0])
        return g

    def batch_train_q(self):
        self.trainer.prep_training()
        self.train_infos = []
        graphed = self._graphed_update()
        if graphed is not None:
            self.train_infos.append(graphed(self.beta_anneal.eval(self.total_train_steps)))
            return
        for p_id in self.policy_ids:
            if self.use_per:
                beta = self.beta_anneal.eval(self.total_train_steps)
                sample = self.buffer.sample(self.batch_size, beta, p_id)
            else:
                sample = self.buffer.sample(self.batch_size)
            train_info, new_priorities, idxes = self.trainer.train_policy_on_batch(sample)
            if self.use_per:
                self.buffer.update_priorities(idxes, new_priorities, p_id)
            self.train_infos.append(train_info)
        if self.use_soft_update:
            self.trainer.soft_target_updates()
        elif (self.num_episodes_collected - self.last_hard_update_episode) / self.hard_update_interval_episode >= 1:
            self.trainer.hard_target_updates()
            self.last_hard_update_episode = self.num_episodes_collected

    @torch.no_grad()
    def eval(self):
        """One greedy batch of episodes (not stored), like magym_runner.py:25-37."""
        self.trainer.prep_rollout()
        info = self.collect_rollout(explore=False, training_episode=False, warmup=False)
        self.eval_reward = info["average_episode_rewards"]
        return self.eval_reward

    def log(self):
        """base_runner.py:308-331 (log_env / log_train): one record per call; with ``self.sink`` (a ``MetricSink``) it goes to
        logging / wandb with ``step=total_env_steps`` and the tensors cross to the host in one copy."""
        from ..utils.sinks import MetricSink
        rec = {"total_env_steps": self.total_env_steps, "episodes": self.num_episodes_collected,
               "average_episode_rewards": self.env_infos.get("average_episode_rewards", float("nan")),
               **({"policy_0/" + k: v for k, v in self.train_infos[-1].items()} if self.train_infos else {})}
        sink = getattr(self, "sink", None)
        self.last_log = (sink.log(rec, step=self.total_env_steps, force=True) if sink is not None else None) or MetricSink.to_host(rec)
        for k in list(self.last_log):                        # keep the un-prefixed spellings older callers read
            if k.startswith("policy_0/"):
                self.last_log[k[len("policy_0/"):]] = self.last_log[k]
        return self.last_log
