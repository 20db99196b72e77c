import logging

import pytest

from tf_yarn_b200 import evaluator_metrics
from tf_yarn_b200.evaluator_metrics import EvaluatorMetricsLogger
from tf_yarn_b200.topologies import ContainerKey, ContainerTask

from fakes import FakeClient

METRICS = {"metric1": "metric1 description", "metric2": "metric2 description"}
evaluators = [ContainerTask("eval", 1, 1), ContainerTask("eval", 2, 1)]

scenarios = [
    # (kv, last metrics, thresholds, expected log lines)
    ({"eval:1/metric1": b"0.0", "eval:1/metric2": b"13.0", "eval:2/metric1": b"5.0", "eval:2/metric2": b"16.0"},
     None, None,
     ["Statistics for eval:1: metric1 description: 0.0 metric2 description: 13.0",
      "Statistics for eval:2: metric1 description: 5.0 metric2 description: 16.0"]),
    ({"eval:1/metric1": b"0.9", "eval:1/metric2": b"13.0", "eval:2/metric1": b"13.0", "eval:2/metric2": b"26.0"},
     {ContainerKey("eval", 1): {"metric1": 0.9, "metric2": 13.0}, ContainerKey("eval", 2): {"metric1": 5.0, "metric2": 16.0}},
     None,
     ["Statistics for eval:2: metric1 description: 13.0 metric2 description: 26.0"]),
    ({"eval:1/metric1": b"0.9", "eval:1/metric2": b"13.0", "eval:2/metric1": b"13.0", "eval:2/metric2": b"26.0"},
     None, {"metric1": (0.0, 1.0), "metric2": (20.0, None)},
     ["Statistics for eval:1: metric1 description: 0.9", "Statistics for eval:2: metric2 description: 26.0"]),
    ({"eval:1/metric1": b"0.9", "eval:1/metric2": b"13.0", "eval:2/metric1": b"13.0", "eval:2/metric2": b"26.0"},
     None, {"metric1": (None, 1.0)},
     ["Statistics for eval:1: metric1 description: 0.9 metric2 description: 13.0",
      "Statistics for eval:2: metric2 description: 26.0"]),
]


@pytest.mark.parametrize("kv,last,thresholds,expected", scenarios)
def test_log(kv, last, thresholds, expected, monkeypatch, caplog):
    monkeypatch.setattr(evaluator_metrics, "MONITORED_METRICS", METRICS)
    app = FakeClient(kv)
    logger = EvaluatorMetricsLogger(evaluators, app, thresholds)
    if last:
        logger.last_metrics = last
    with caplog.at_level(logging.INFO, logger="tf_yarn_b200.evaluator_metrics"):
        logger.log()
    assert [r.getMessage() for r in caplog.records] == expected
    # state is updated whether or not the value was inside the thresholds
    assert logger.last_metrics[ContainerKey("eval", 2)]["metric2"] == float(kv["eval:2/metric2"])


def test_unknown_threshold_key_warns(monkeypatch):
    monkeypatch.setattr(evaluator_metrics, "MONITORED_METRICS", METRICS)
    with pytest.warns(UserWarning):
        EvaluatorMetricsLogger(evaluators, FakeClient(), {"nope": (0, 1)})
