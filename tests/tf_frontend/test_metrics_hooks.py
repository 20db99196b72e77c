"""Monitoring hooks of the TF front-end, in process (reference: tf_yarn/tensorflow/metrics.py:18-142 and its
tests/tensorflow tests): steps/sec hook, evaluator KV metrics, hook injection into an Experiment."""
import pickle
from unittest import mock

import pytest

from fakes import FakeClient
from tf_yarn_b200 import estimator as est
from tf_yarn_b200.estimator.hooks import SessionRunContext, SessionRunValues
from tf_yarn_b200.tensorflow import Experiment, KerasExperiment
from tf_yarn_b200.tensorflow import metrics as tfm


def test_step_per_second_hook_logs_on_the_chief_only(monkeypatch):
    logged = []
    monkeypatch.setattr(tfm.mlflow, "log_metric", lambda key, value, step=None: logged.append((key, value, step)))
    monkeypatch.setenv("TFY_N_TRY", "2")
    hook = tfm.StepPerSecondHook(every_n_steps=10)
    monkeypatch.setenv("TFY_TASK_KEY", "worker:0")
    hook._log_and_record(10, 2.0, 40)
    assert logged == [] and hook.last_steps_per_sec == 5.0
    monkeypatch.setenv("TFY_TASK_KEY", "chief:0")
    hook._log_and_record(30, 2.0, 70)
    assert logged == [("steps_per_sec_2", 15.0, 70)]
    with pytest.raises(ValueError):
        tfm.StepPerSecondHook(every_n_steps=None, every_n_secs=None)


def test_eval_monitor_hook_publishes_the_four_evaluator_metrics(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "evaluator:0")
    client = FakeClient()
    hook = tfm.EvalMonitorHook(client)
    ctx = SessionRunContext(None, 120)
    for _ in range(3):
        args = hook.before_run(ctx)
        assert args is not None
        hook.after_run(ctx, SessionRunValues(results=120))
    kv = client.kv
    assert kv["evaluator:0/nb_eval_steps"] == b"3" and kv["evaluator:0/last_training_step"] == b"120"
    assert 0.0 <= float(kv["evaluator:0/awake_time_ratio"]) <= 1.0
    assert float(kv["evaluator:0/eval_step_mean_duration"]) >= 0.0
    clone = pickle.loads(pickle.dumps(hook))                      # experiments are shipped pickled: no live client inside
    assert clone._client is None and clone.step_counter == 3


def _experiment(hooks=()):
    e = est.LinearClassifier([est.feature_column.numeric_column("x", shape=(2,))], n_classes=2,
                             config=est.RunConfig(log_step_count_steps=7))
    fn = lambda: iter(())   # noqa: E731
    return Experiment(e, est.TrainSpec(fn, max_steps=1, hooks=list(hooks)), est.EvalSpec(fn, steps=1))


def test_monitor_hooks_are_injected_once():
    exp = tfm._add_monitor_to_experiment(_experiment())
    assert [type(h).__name__ for h in exp.train_spec.hooks] == ["StepPerSecondHook"]
    assert exp.train_spec.hooks[0]._every_steps == 7
    assert type(exp.eval_spec.hooks[0]).__name__ == "EvalMonitorHook"
    again = tfm._add_monitor_to_experiment(_experiment([tfm.StepPerSecondHook(every_n_steps=3)]))
    assert len(again.train_spec.hooks) == 1 and again.train_spec.hooks[0]._every_steps == 3      # the user's hook wins
    keras_exp = KerasExperiment(mock.Mock(), "/tmp/x", {})
    assert tfm._add_monitor_to_experiment(keras_exp) is keras_exp
    with pytest.raises(ValueError):
        tfm._add_monitor_to_experiment(object())
