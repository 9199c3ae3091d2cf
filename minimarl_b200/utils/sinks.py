"""Logging sinks fed from device-side accumulators.

The reference logs from Python floats every episode: ``logging.info`` lines and ``wandb.log`` dicts in the top-level loops
(``vdn/main.py:180-196``, ``qmix/main.py``), ``wandb.log(train_infos)`` in the MAPPO runner
(``mappo/runner/shared/magym_runner.py:104-105``), ``print`` + ``wandb.log(..., step=total_env_steps)`` in the off-policy
runner (``offpolicy/runner/shared/base_runner.py:308-331``).  Here the values are device tensors (episodic returns summed
by the rollout kernels, losses of graph-replayed updates); ``MetricSink.log`` takes a dict of floats / 0-dim tensors, brings
ALL tensor values to the host with one stacked copy (one sync per log call instead of one per value), and forwards the
record to

* the ``logging`` module, with the reference's line formats (``vdn_line`` / ``test_line``),
* ``wandb`` when ``use_wandb`` and the package is importable (``mode="offline"`` unless told otherwise: no network here),
* ``history`` (a list of dicts), which is what the tests and ``bench.py`` read.

``every`` thins the stream: with thousands of lockstep envs an "episode" of the loop is thousands of reference episodes,
and a sync per episode would serialise the device.
"""
import logging
import os
import time

import torch


class MetricSink:
    def __init__(self, use_wandb=False, use_logging=True, every=1, wandb_kwargs=None, log_dir=None, experiment_name="minimarl_b200",
                 logger=None):
        self.every = max(1, int(every))
        self.history = []
        self._n = 0
        self._wandb = None
        self.logger = logger
        if use_logging and logger is None:
            self.logger = logging.getLogger(f"minimarl_b200.{experiment_name}")
            self.logger.setLevel(logging.INFO)
            if log_dir is not None and not self.logger.handlers:                     # the reference's set_logging (vdn/_utils.py:55-60)
                os.makedirs(log_dir, exist_ok=True)
                h = logging.FileHandler(os.path.join(log_dir, f"{experiment_name}-{int(time.time())}.log"))
                h.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s - %(message)s"))
                self.logger.addHandler(h)
        if use_wandb:
            try:
                import wandb
                kw = dict(mode=os.environ.get("WANDB_MODE", "offline"), reinit=True, name=f"{experiment_name}-{int(time.time())}")
                kw.update(wandb_kwargs or {})
                self._run = wandb.init(**kw)
                self._wandb = wandb
            except Exception as e:                                  # a missing / unconfigured wandb must not stop training
                if self.logger:
                    self.logger.warning("wandb disabled: %s", e)

    @staticmethod
    def to_host(record):
        """dict of numbers / tensors -> the same dict with every tensor value as a Python float, ONE device-to-host copy for all."""
        keys = [k for k, v in record.items() if torch.is_tensor(v)]
        out = dict(record)
        if keys:
            flat = torch.stack([record[k].detach().reshape(()).to(torch.float64) for k in keys]).tolist()
            out.update(zip(keys, flat))
        return out

    def log(self, record, step=None, line=None, force=False):
        """Returns the host record when it was emitted, else None (thinned by ``every``; tensors stay on the device)."""
        self._n += 1
        if not force and (self._n - 1) % self.every:
            return None
        rec = self.to_host(record)
        self.history.append(rec)
        if self.logger is not None:
            self.logger.info(line(rec) if line else " | ".join(f"{k}: {v}" for k, v in rec.items()))
        if self._wandb is not None:
            self._wandb.log(rec, step=step)
        return rec

    def close(self):
        if self._wandb is not None:
            self._wandb.finish()
            self._wandb = None


def vdn_line(max_episodes):
    """The per-episode line of vdn/main.py:180 / qmix/main.py."""
    def fmt(r):
        return (f"{int(r['episode']) + 1:<8}/{max_episodes:<10} episodes | train score: {r['train_score']:<6.2f} | "
                f"train loss: {r['train_loss']:<7.2f} | eps: {r['eps']:<5.4f} | alpha: {r['alpha']:<6.4f} | beta: {r['beta']:<6.4f}")
    return fmt


def test_line(r):
    """vdn/main.py:188."""
    return f"avg test score: {r['test-score']:<6.2f} | avg test loss: {r['test-loss']:<6.2f}"
