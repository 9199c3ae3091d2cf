"""The training loop of ``vdn/main.py`` (:80-198) for N lockstep Checkers boards (full observations, D = 94).

Per step: ``behavior_network.sample_action`` (one epsilon draw per env) -> lockstep env kernel ->
``target_network`` forward -> ``cal_td_error`` for every env -> every ``chunk_size`` steps the chunks of all
envs go into the device chunk-PER; after each batch of episodes ``Target_Dqn.train`` and the periodic hard
target copy, with the reference's epsilon / alpha / beta schedules.
"""
import torch

from ..algorithms.toplevel import Prioritized_Experience_Replay, Q_Net, Target_Dqn, cal_td_error
from ..envs.batched import BatchedCheckers


def run_vdn(args, n_envs, episodes, device="cuda", log=None):
    dev = torch.device(device)
    env = BatchedCheckers(n_envs, full_observable=True, max_steps=args.max_step, step_cost=args.step_cost, device=dev)
    beh = Q_Net(env.observation_space, env.action_space, args).to(dev)
    tgt = Q_Net(env.observation_space, env.action_space, args).to(dev)
    tgt.load_state_dict(beh.state_dict())
    per = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=env.obs_dim, device=dev)
    opt = torch.optim.Adam(beh.parameters(), lr=args.lr)
    trainer = Target_Dqn(per, beh, tgt, args, dev)
    L = args.chunk_size
    history = []
    for episode in range(episodes):
        epsilon = max(args.min_epsilon, args.max_epsilon - (args.max_epsilon - args.min_epsilon) *
                      (episode / args.epsilon_anneal_episode))
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
                if (step + 1) % L == 0:
                    stack = {k: torch.stack(v, 1) for k, v in chunk.items()}
                    for e in range(n_envs):              # one chunk per env, in env order (the reference's add order)
                        per.collect_sample([stack["s"][e], stack["a"][e], stack["r"][e], stack["n"][e], stack["d"][e]],
                                           td_sum[e].item(), warm_up=False)
                    chunk = {k: [] for k in "sarnd"}
                    td_sum.zero_()
                state, hidden = next_state, next_hidden
                if not bool(alive.any()):
                    break
        loss = trainer.train(target_network=tgt, optimizer=opt, epsilon=epsilon) if per.sum_tree.n_data >= args.batch_size \
            else torch.tensor(float("nan"))
        if episode % args.update_target_interval == 0:
            tgt.load_state_dict(beh.state_dict())
        rec = {"episode": episode, "train_score": score.mean().item(), "train_loss": float(loss), "eps": epsilon,
               "alpha": per.alpha, "beta": per.beta}
        history.append(rec)
        if log:
            log(rec)
    return history, beh
