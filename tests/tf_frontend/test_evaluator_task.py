import os
from unittest import mock

import torch

from tf_yarn_b200 import data, keras
from tf_yarn_b200 import estimator as est
from tf_yarn_b200.tensorflow import Experiment, KerasExperiment
from tf_yarn_b200.tensorflow.tasks import evaluator_task

from fakes import FakeClient

fc = est.feature_column


def _experiment(model_dir, max_steps=100):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(256, 4, generator=g)
    y = (x[:, 0] > 0).long()
    train = lambda: data.Dataset.from_tensor_slices(({"x": x}, y)).batch(32).repeat()  # noqa: E731
    evalf = lambda: data.Dataset.from_tensor_slices(({"x": x}, y)).batch(64)  # noqa: E731
    e = est.LinearClassifier([fc.numeric_column("x", shape=(4,))], model_dir=str(model_dir), n_classes=2,
                             optimizer=lambda: keras.optimizers.Adam(0.05),
                             config=est.RunConfig(save_checkpoints_steps=50, log_step_count_steps=None))
    return Experiment(e, est.TrainSpec(train, max_steps=max_steps),
                      est.EvalSpec(evalf, steps=2, start_delay_secs=0, throttle_secs=0))


def test_evaluate_only_not_yet_evaluated_checkpoints(tmp_path):
    exp = _experiment(tmp_path)
    exp.estimator.train(exp.train_spec.input_fn, max_steps=100)
    # pretend step 0 and 50 were evaluated by a previous evaluator incarnation
    exp.estimator.evaluate(exp.eval_spec.input_fn, steps=1, checkpoint_path=str(tmp_path / "model.ckpt-0"))
    exp.estimator.evaluate(exp.eval_spec.input_fn, steps=1, checkpoint_path=str(tmp_path / "model.ckpt-50"))
    assert evaluator_task.get_initial_evaluated_checkpoints(str(tmp_path / "eval")) == {0, 50}
    with mock.patch.object(exp.estimator, "evaluate", wraps=exp.estimator.evaluate) as spy:
        evaluator_task.evaluate(exp, timeout_in_secs=30)
    assert [os.path.basename(c.kwargs["checkpoint_path"]) for c in spy.call_args_list] == ["model.ckpt-100"]


def test_evaluate_returns_immediately_when_max_steps_already_evaluated(tmp_path):
    exp = _experiment(tmp_path, max_steps=50)
    exp.estimator.train(exp.train_spec.input_fn, max_steps=50)
    exp.estimator.evaluate(exp.eval_spec.input_fn, steps=1)
    with mock.patch.object(exp.estimator, "evaluate") as spy:
        assert evaluator_task.evaluate(exp, timeout_in_secs=30) is None
    spy.assert_not_called()


def test_get_all_checkpoints_and_step(tmp_path):
    exp = _experiment(tmp_path)
    exp.estimator.train(exp.train_spec.input_fn, max_steps=60)
    names = [os.path.basename(p) for p in evaluator_task._get_all_checkpoints(str(tmp_path))]
    assert names == ["model.ckpt-0", "model.ckpt-50", "model.ckpt-60"]
    assert evaluator_task._get_step("x/model.ckpt-123") == 123
    assert evaluator_task._get_step("dir/checkpoint-7.ckpt") == 7


def test_keras_evaluate_picks_up_checkpoints_and_stops(tmp_path, monkeypatch):
    monkeypatch.setattr(evaluator_task, "KERAS_POLL_SECS", 0.05)
    monkeypatch.setenv("TFY_TASK_KEY", "evaluator:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 4, generator=g)
    y = (x[:, 0] > 0).long()
    m = keras.Sequential([keras.layers.Dense(2, input_shape=(4,))])
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer="sgd", metrics=["accuracy"])
    m.build()
    m.save(str(tmp_path / "checkpoint-1.ckpt"))
    m.save(str(tmp_path / "checkpoint-2.ckpt"))
    exp = KerasExperiment(m, str(tmp_path), {}, validation_data_fn=lambda: data.Dataset.from_tensor_slices((x, y)).batch(32))
    client = FakeClient()
    evaluator_task.keras_evaluate(exp, stop_cond=lambda: True, timeout_in_secs=10, client=client)
    from tf_yarn_b200.estimator import summary
    scalars = summary.read_scalars(str(tmp_path / "eval"))
    assert sorted(set(scalars["step"])) == [1, 2] and "accuracy" in scalars["name"]
    assert client.kv["evaluator:0/nb_eval_steps"] == b"2"
