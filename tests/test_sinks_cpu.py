"""CPU: the logging sinks (minimarl_b200/utils/sinks.py) -- one host copy per record, the reference's line formats,
thinning, and a wandb failure never stops the loop."""
import logging

import torch

from minimarl_b200.utils.sinks import MetricSink, vdn_line
from minimarl_b200.utils.sinks import test_line as eval_line


def test_to_host_and_reference_line_formats(caplog):
    rec = {"episode": 4, "train_score": torch.tensor(8.0), "train_loss": torch.tensor(1.25), "eps": 0.7998, "alpha": 0.6,
           "beta": torch.tensor(0.4, dtype=torch.float64)}
    host = MetricSink.to_host(rec)
    assert host == {"episode": 4, "train_score": 8.0, "train_loss": 1.25, "eps": 0.7998, "alpha": 0.6, "beta": 0.4}
    # the f-string of vdn/main.py:180
    episode, max_episodes, train_score, train_loss, epsilon, alpha, beta = 4, 30000, 8.0, 1.25, 0.7998, 0.6, 0.4
    want = (f"{(episode+1):<8}/{max_episodes:<10} episodes | train score: {train_score:<6.2f} | train loss: {train_loss:<7.2f} | "
            f"eps: {epsilon:<5.4f} | alpha: {alpha:<6.4f} | beta: {beta:<6.4f}")
    assert vdn_line(30000)(host) == want
    assert eval_line({"test-score": 7.5, "test-loss": 0.25}) == f"avg test score: {7.5:<6.2f} | avg test loss: {0.25:<6.2f}"
    logger = logging.getLogger("sink-test")
    sink = MetricSink(use_wandb=False, logger=logger, every=2)
    with caplog.at_level(logging.INFO, logger="sink-test"):
        assert sink.log(rec, line=vdn_line(30000)) is not None          # 1st record emitted
        assert sink.log(rec, line=vdn_line(30000)) is None              # 2nd thinned (no host copy at all)
        assert sink.log({"test-score": 1.0, "test-loss": 2.0}, line=eval_line, force=True) is not None
    assert len(sink.history) == 2 and want in caplog.text and "avg test score" in caplog.text


def test_wandb_failure_is_not_fatal(monkeypatch):
    import sys
    import types
    fake = types.ModuleType("wandb")

    def boom(**kw):
        raise RuntimeError("no network")
    fake.init = boom
    monkeypatch.setitem(sys.modules, "wandb", fake)
    sink = MetricSink(use_wandb=True, logger=logging.getLogger("sink-test2"))
    assert sink._wandb is None
    assert sink.log({"x": torch.tensor(1.0)}) == {"x": 1.0}
    sink.close()
