"""mini-Estimator on CPU: train/evaluate/predict, checkpoints, hooks, continuous eval, exporters."""
import os

import pytest
import torch

from tf_yarn_b200 import data, keras
from tf_yarn_b200 import estimator as est
from tf_yarn_b200.estimator import checkpoint as ckpt

fc = est.feature_column


def _data(n=1024, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 6, generator=g)
    w = torch.randn(6, 3, generator=g)
    return x, (x @ w).argmax(1)


def _input_fns():
    x, y = _data()
    train = lambda: data.Dataset.from_tensor_slices(({"x": x}, y)).shuffle(500, seed=1).batch(64).repeat()  # noqa: E731
    evalf = lambda: data.Dataset.from_tensor_slices(({"x": x}, y)).batch(256)  # noqa: E731
    return train, evalf


def _classifier(model_dir, **cfg):
    return est.LinearClassifier([fc.numeric_column("x", shape=(6,))], model_dir=str(model_dir), n_classes=3,
                                optimizer=lambda: keras.optimizers.Adam(0.05),
                                config=est.RunConfig(save_checkpoints_steps=50, log_step_count_steps=None, **cfg))


def test_train_evaluate_predict_and_checkpoint_index(tmp_path):
    train, evalf = _input_fns()
    e = _classifier(tmp_path)
    e.train(train, max_steps=120)
    state = est.get_checkpoint_state(str(tmp_path))
    assert [ckpt.step_of(p) for p in state.all_model_checkpoint_paths] == [0, 50, 100, 120]
    assert est.latest_checkpoint(str(tmp_path)).endswith("model.ckpt-120")
    res = e.evaluate(evalf)
    assert res["global_step"] == 120 and res["accuracy"] > 0.85
    first = next(iter(e.predict(evalf)))
    assert set(first) == {"logits", "probabilities", "class_ids"}
    # evaluation of an older checkpoint reports that checkpoint's step
    old = e.evaluate(evalf, checkpoint_path=state.all_model_checkpoint_paths[1], name="old")
    assert old["global_step"] == 50
    assert os.path.isdir(tmp_path / "eval_old")


def test_training_resumes_from_latest_checkpoint(tmp_path):
    train, _ = _input_fns()
    _classifier(tmp_path).train(train, max_steps=60)
    e2 = _classifier(tmp_path)
    e2.train(train, max_steps=100)
    assert e2.get_global_step() == 100
    e3 = _classifier(tmp_path)
    e3.train(train, max_steps=100)           # already there: no further step
    assert e3.get_global_step() == 100


def test_hooks_receive_global_step_and_can_stop(tmp_path):
    train, _ = _input_fns()
    seen = []

    class Spy(est.SessionRunHook):
        def before_run(self, ctx):
            return est.SessionRunArgs(est.get_global_step())

        def after_run(self, ctx, values):
            seen.append(values.results)
    e = _classifier(tmp_path)
    e.train(train, hooks=[Spy(), est.StopAtStepHook(last_step=7)], max_steps=1000)
    assert seen == list(range(1, 8)) and e.get_global_step() == 7


def test_step_counter_hook_reports_rate():
    hook = est.StepCounterHook(every_n_steps=2)
    ctx = est.SessionRunContext()
    for step in range(1, 6):
        hook.after_run(ctx, est.SessionRunValues(results=step))
    assert hook.last_steps_per_sec is not None and hook.last_steps_per_sec > 0
    with pytest.raises(ValueError):
        est.StepCounterHook(every_n_steps=None, every_n_secs=None)


def test_train_and_evaluate_local_with_exporter(tmp_path, monkeypatch):
    monkeypatch.delenv("TF_CONFIG", raising=False)
    train, evalf = _input_fns()
    e = _classifier(tmp_path)
    res, _ = est.train_and_evaluate(e, est.TrainSpec(train, max_steps=80),
                                    est.EvalSpec(evalf, steps=None, exporters=est.FinalExporter("final")))
    assert res["global_step"] == 80
    exports = os.listdir(tmp_path / "export" / "final")
    assert len(exports) == 1 and os.path.exists(tmp_path / "export" / "final" / exports[0] / "saved_model.pt")


def test_continuous_eval_evaluates_only_new_checkpoints_and_stops_at_max_steps(tmp_path):
    train, evalf = _input_fns()
    e = _classifier(tmp_path)
    e.train(train, max_steps=100)
    evaluated = []

    class Rec(est.Exporter):
        def export(self, estimator, export_path, checkpoint_path, eval_result, is_final):
            evaluated.append((ckpt.step_of(checkpoint_path), is_final))
    res = est.continuous_eval(e, est.TrainSpec(train, max_steps=100),
                              est.EvalSpec(evalf, steps=2, exporters=[Rec("rec")], start_delay_secs=0, throttle_secs=0),
                              timeout_secs=30, evaluated_steps={0})
    assert [s for s, _ in evaluated] == [50, 100] and evaluated[-1][1] is True
    assert res["global_step"] == 100


def test_dnn_and_wide_deep_estimators_learn(tmp_path):
    g = torch.Generator().manual_seed(3)
    n = 1024
    num = torch.randn(n, 4, generator=g)
    cat = torch.randint(0, 50, (n, 1), generator=g)
    y = ((num[:, 0] > 0) ^ (cat[:, 0] % 2 == 0)).long()
    feats = {"num": num, "cat": cat}
    train = lambda: data.Dataset.from_tensor_slices((feats, y)).shuffle(500, seed=0).batch(64).repeat()  # noqa: E731
    evalf = lambda: data.Dataset.from_tensor_slices((feats, y)).batch(256)  # noqa: E731
    cat_col = fc.categorical_column_with_identity("cat", 50)
    e = est.DNNLinearCombinedClassifier(
        model_dir=str(tmp_path), linear_feature_columns=[cat_col],
        dnn_feature_columns=[fc.numeric_column("num", shape=(4,)), fc.embedding_column(cat_col, 8)],
        dnn_hidden_units=[32, 16], dnn_optimizer=lambda: keras.optimizers.Adam(0.01),
        config=est.RunConfig(save_checkpoints_steps=None, save_checkpoints_secs=None, log_step_count_steps=None))
    e.train(train, max_steps=400)
    assert e.evaluate(evalf)["accuracy"] > 0.85


def test_model_to_estimator(tmp_path):
    x, y = _data()
    m = keras.Sequential([keras.layers.Dense(16, activation="relu", input_shape=(6,)), keras.layers.Dense(3)])
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer=keras.optimizers.Adam(0.02),
              metrics=["accuracy"])
    e = est.model_to_estimator(m, model_dir=str(tmp_path),
                               config=est.RunConfig(save_checkpoints_steps=None, save_checkpoints_secs=None,
                                                    log_step_count_steps=None))
    train = lambda: data.Dataset.from_tensor_slices(({"features": x}, y)).batch(64).repeat()  # noqa: E731
    e.train(train, max_steps=200)
    assert e.evaluate(lambda: data.Dataset.from_tensor_slices(({"features": x}, y)).batch(256))["accuracy"] > 0.8


def test_run_config_replace_and_cluster_info(monkeypatch):
    cfg = est.RunConfig(model_dir="/m", session_config=est.SessionConfig(device_filters=["/job:ps"]))
    assert cfg.replace(model_dir=None, save_summary_steps=None).model_dir is None and cfg.model_dir == "/m"
    with pytest.raises(ValueError):
        cfg.replace(nope=1)
    monkeypatch.setenv("TF_CONFIG", '{"cluster": {"chief": ["a:1"], "worker": ["b:2", "c:3"], "ps": ["d:4"]}, '
                                    '"task": {"type": "worker", "index": 1}}')
    info = est.ClusterInfo.from_env()
    assert (info.task_type, info.task_id, info.has_ps, info.is_chief) == ("worker", 1, True, False)
    assert info.trainers() == ["chief:0", "worker:0", "worker:1"]
    assert cfg.num_ps_replicas == 1 and cfg.num_worker_replicas == 3


def test_ftrl_optimizer_learns_and_per_tower_optimizers_are_resolved():
    """FTRL (TF's default for linear models) trains a logistic regression; DNNLinearCombinedClassifier routes the
    wide tower ("linear.*") to FTRL and the deep tower to Adagrad."""
    import torch
    from tf_yarn_b200 import estimator as est
    from tf_yarn_b200 import keras
    from tf_yarn_b200.estimator import feature_column as fc
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.zeros(10))
    opt = keras.optimizers.Ftrl(0.1, l1_regularization_strength=0.01).to_torch([w])
    x = torch.randn(256, 10)
    y = (x[:, 0] - 2 * x[:, 3] > 0).float()
    for _ in range(200):
        opt.zero_grad()
        loss = torch.nn.functional.binary_cross_entropy_with_logits(x @ w, y)
        loss.backward()
        opt.step()
    assert loss.item() < 0.35, loss.item()              # from ln 2 = 0.69
    wd = w.detach()
    assert wd[0] > 0.5 and wd[3] < -1.0 and wd.abs()[[1, 2, 4, 5, 6, 7, 8, 9]].max() < 0.5

    num = fc.numeric_column("x", shape=(4,))
    cat = fc.categorical_column_with_hash_bucket("c", 50)
    e = est.DNNLinearCombinedClassifier(linear_feature_columns=[num, cat],
                                        dnn_feature_columns=[num, fc.embedding_column(cat, 8)],
                                        dnn_hidden_units=[16], config=est.RunConfig(save_checkpoints_steps=None,
                                                                                    save_checkpoints_secs=None))

    def input_fn():
        from tf_yarn_b200.data import Dataset
        g = torch.Generator().manual_seed(1)
        batches = []
        for _ in range(8):
            xs = torch.randn(32, 4, generator=g)
            cs = torch.randint(0, 50, (32, 1), generator=g)
            batches.append(({"x": xs, "c": cs}, (xs[:, 0] > 0).long()))
        return Dataset(lambda: iter(batches), len(batches))
    e.train(input_fn, steps=8)
    kinds = {n: d.to_spec().kind for n, d in e._opt_by_name.items()}
    assert all(k == "ftrl" for n, k in kinds.items() if n.startswith("linear."))
    assert all(k == "adagrad" for n, k in kinds.items() if not n.startswith("linear."))
    assert len({k for k in kinds.values()}) == 2


def test_ps_layout_carries_per_variable_optimizers():
    from tf_yarn_b200.estimator import ps
    lay = ps.Layout([("linear.w", [10]), ("dnn.w", [4, 4])], 2, "adagrad", {"lr": 0.1, "p1": 0, "p2": 0, "eps": 1e-7,
                    "wd": 0, "init_s1": 0.1, "flags": 0}, kinds=["ftrl", "adagrad"],
                    hypers=[{"lr": 0.05, "p1": 0.01, "p2": 0, "eps": 0, "wd": 0, "init_s1": 0.1, "flags": 0}] * 2)
    again = ps.Layout.from_json(lay.to_json())
    assert again.kinds == ["ftrl", "adagrad"] and again.var_slots == [2, 1]
    assert again.shard_elems[0] == 16 * 3 and again.shard_elems[1] == 16 * 2


def test_latest_and_best_exporters_with_a_recording_estimator(tmp_path):
    """LatestExporter keeps the N most recent numeric export directories; BestExporter exports only improvements."""

    class FakeEstimator:
        def __init__(self):
            self.n = 0

        def export_saved_model(self, export_path, checkpoint_path=None):
            self.n += 1
            out = os.path.join(export_path, str(1000 + self.n))
            os.makedirs(out, exist_ok=True)
            return out

    fake = FakeEstimator()
    latest = est.LatestExporter("latest", exports_to_keep=2)
    path = str(tmp_path / "export" / "latest")
    os.makedirs(path)
    for i in range(4):
        latest.export(fake, path, f"model.ckpt-{i}", {"loss": 1.0}, False)
    assert sorted(os.listdir(path)) == ["1003", "1004"] and latest.name == "latest"

    best = est.BestExporter("best")
    bpath = str(tmp_path / "export" / "best")
    os.makedirs(bpath)
    outs = [best.export(fake, bpath, "c", {"loss": v}, False) for v in (0.9, 1.2, 0.5, 0.5)]
    assert [o is not None for o in outs] == [True, False, True, False]

    final = est.FinalExporter("final")
    assert final.export(fake, bpath, "c", {"loss": 1.0}, False) is None
    assert final.export(fake, bpath, "c", {"loss": 1.0}, True) is not None


def test_per_tower_optimizer_state_survives_a_checkpoint(tmp_path):
    """FTRL (wide) + Adagrad (deep) state is saved with the checkpoint and restored into a fresh Estimator, which
    resumes at the saved global step (several torch optimizers behind one interface: _MultiOptimizer)."""
    import torch
    from tf_yarn_b200 import estimator as est
    from tf_yarn_b200.estimator import feature_column as fc

    def make():
        num = fc.numeric_column("x", shape=(4,))
        cat = fc.categorical_column_with_hash_bucket("c", 50)
        return est.DNNLinearCombinedClassifier(
            linear_feature_columns=[num, cat], dnn_feature_columns=[num, fc.embedding_column(cat, 8)],
            dnn_hidden_units=[16], model_dir=str(tmp_path),
            config=est.RunConfig(save_checkpoints_steps=4, tf_random_seed=3))

    def input_fn():
        from tf_yarn_b200.data import Dataset
        g = torch.Generator().manual_seed(1)
        batches = []
        for _ in range(8):
            xs = torch.randn(32, 4, generator=g)
            cs = torch.randint(0, 50, (32, 1), generator=g)
            batches.append(({"x": xs, "c": cs}, (xs[:, 0] > 0).long()))
        return Dataset(lambda: iter(batches), len(batches)).repeat()

    first = make()
    first.train(input_fn, max_steps=8)
    saved = first._optimizer.state_dict()
    assert "multi" in saved and len(saved["multi"]) == 2 and len(first._optimizer.param_groups) >= 2

    second = make()
    second.train(input_fn, max_steps=9)                       # restores step 8, takes ONE more step
    assert second.get_global_step() == 9

    third = make()
    third.train(input_fn, max_steps=9)                        # nothing left to do: state is exactly what step 9 saved
    restored = third._optimizer.state_dict()
    mid = second._optimizer.state_dict()
    for a, b in zip(restored["multi"], mid["multi"]):
        for k in a["state"]:
            for name, v in a["state"][k].items():
                if torch.is_tensor(v):
                    assert torch.allclose(v, b["state"][k][name]), name
    # accumulators grew past their initial value: the restored state is the trained one, not a fresh optimizer
    flat = [v for st in restored["multi"] for s in st["state"].values() for v in s.values() if torch.is_tensor(v)]
    assert flat and any(float(v.abs().max()) > 0.1 for v in flat)


def test_continuous_eval_stops_on_idle_timeout_and_on_stop_condition(tmp_path):
    """No checkpoint ever appears: the evaluator gives up after its idle timeout, or as soon as the stop condition
    (every trainer reported `stop`) holds (reference: tf_yarn/tensorflow/tasks/evaluator_task.py:18-25,83-127)."""
    import time
    from tf_yarn_b200.estimator import training

    class Idle:
        model_dir = str(tmp_path)

        def evaluate(self, *a, **k):
            raise AssertionError("nothing to evaluate")

    spec = est.EvalSpec(lambda: iter(()), steps=1, start_delay_secs=0, throttle_secs=0)
    t0 = time.time()
    assert training.continuous_eval(Idle(), est.TrainSpec(lambda: iter(()), max_steps=10), spec, timeout_secs=0.3) is None
    assert 0.25 < time.time() - t0 < 5.0
    calls = []
    assert training.continuous_eval(Idle(), est.TrainSpec(lambda: iter(()), max_steps=10), spec, timeout_secs=60,
                                    stop_cond=lambda: calls.append(1) or len(calls) >= 3) is None
    assert len(calls) == 3


def test_exports_in_the_same_second_get_distinct_directories_and_predict_modes(tmp_path):
    import torch
    e = _classifier(tmp_path / "m") if "_classifier" in globals() else None
    if e is None:
        from tf_yarn_b200.estimator import feature_column as fc
        e = est.LinearClassifier([fc.numeric_column("x", shape=(6,))], model_dir=str(tmp_path / "m"), n_classes=3,
                                 config=est.RunConfig(save_checkpoints_steps=5))
    g = torch.Generator().manual_seed(0)
    xs = torch.randn(40, 6, generator=g)
    ys = torch.randint(0, 3, (40,), generator=g)

    def input_fn():
        from tf_yarn_b200.data import Dataset
        return Dataset.from_tensor_slices(({"x": xs}, ys)).batch(8)
    e.train(input_fn, max_steps=5)
    base = str(tmp_path / "export")
    outs = [e.export_saved_model(base) for _ in range(3)]
    assert len(set(outs)) == 3 and sorted(outs) == outs
    assert all(os.path.exists(os.path.join(o, "saved_model.pt")) for o in outs)
    single = list(e.predict(input_fn))
    assert len(single) == 40 and single[0]["probabilities"].shape == (3,)
    batched = list(e.predict(input_fn, yield_single_examples=False))
    assert len(batched) == 5 and tuple(batched[0]["probabilities"].shape) == (8, 3)


def test_a_checkpoint_pruned_before_it_was_loaded_is_skipped_not_misreported(tmp_path):
    """The chief prunes old checkpoints (keep_checkpoint_max) while the evaluator works through its list: a checkpoint
    that vanished must raise from evaluate(checkpoint_path=...) and be skipped by the evaluator loop, never be
    'evaluated' with the weights of another step."""
    import torch
    from tf_yarn_b200.estimator import checkpoint as ckpt
    from tf_yarn_b200.estimator import feature_column as fc
    from tf_yarn_b200.estimator import training
    e = est.LinearClassifier([fc.numeric_column("x", shape=(3,))], model_dir=str(tmp_path), n_classes=2,
                             config=est.RunConfig(save_checkpoints_steps=2, keep_checkpoint_max=10))
    g = torch.Generator().manual_seed(0)
    xs, ys = torch.randn(32, 3, generator=g), torch.randint(0, 2, (32,), generator=g)

    def input_fn():
        from tf_yarn_b200.data import Dataset
        return Dataset.from_tensor_slices(({"x": xs}, ys)).batch(8).repeat()
    e.train(input_fn, max_steps=6)
    paths = ckpt.get_checkpoint_state(str(tmp_path)).all_model_checkpoint_paths
    assert [ckpt.step_of(p) for p in paths] == [0, 2, 4, 6]
    with pytest.raises(FileNotFoundError):
        e.evaluate(input_fn, steps=1, checkpoint_path=str(tmp_path / "model.ckpt-999"))

    seen = []
    real_evaluate = e.evaluate

    def evaluate_and_prune(input_fn, steps=None, hooks=None, name=None, checkpoint_path=None):
        if ckpt.step_of(checkpoint_path) == 0:           # while step 0 is evaluated the chief prunes step 2
            os.remove(str(tmp_path / "model.ckpt-2"))
        res = real_evaluate(input_fn, steps=steps, hooks=hooks, name=name, checkpoint_path=checkpoint_path)
        seen.append(int(res["global_step"]))
        return res
    e.evaluate = evaluate_and_prune
    last = training.continuous_eval(e, est.TrainSpec(input_fn, max_steps=6),
                                    est.EvalSpec(input_fn, steps=1, start_delay_secs=0, throttle_secs=0), timeout_secs=30)
    assert seen == [0, 4, 6] and int(last["global_step"]) == 6


def test_logging_and_nan_hooks_take_tfs_signatures(tmp_path, caplog):
    """tf.estimator.LoggingTensorHook(tensors, every_n_iter=...) / NanTensorHook(loss, fail_on_nan_loss=...) as user
    code writes them; StopAtStepHook(num_steps=...) counts from the first step of THIS train() call."""
    import torch
    from tf_yarn_b200.estimator import feature_column as fc
    xs = torch.randn(64, 3)
    ys = (xs[:, 0] > 0).long()

    def input_fn():
        from tf_yarn_b200.data import Dataset
        return Dataset.from_tensor_slices(({"x": xs}, ys)).batch(8).repeat()

    def make(sub):
        return est.LinearClassifier([fc.numeric_column("x", shape=(3,))], model_dir=str(tmp_path / sub), n_classes=2,
                                    config=est.RunConfig(save_checkpoints_steps=None, save_checkpoints_secs=None))
    e = make("a")
    with caplog.at_level("INFO"):
        e.train(input_fn, hooks=[est.LoggingTensorHook({"loss": "loss"}, every_n_iter=4, at_end=True,
                                                       formatter=lambda v: f"L{v['step']}={v['loss']:.3f}"),
                                 est.StopAtStepHook(num_steps=9)], max_steps=1000)
    assert e.get_global_step() == 9
    logged = [r.message for r in caplog.records if r.name.endswith("estimator.hooks") and r.message.startswith("L")]
    assert [m.split("=")[0] for m in logged] == ["L4", "L8", "L9"]              # every 4 steps + at the end
    e.train(input_fn, hooks=[est.StopAtStepHook(num_steps=3)], max_steps=1000)  # 3 MORE steps
    assert e.get_global_step() == 12
    assert est.LoggingTensorHook(50).every_n_iter == 50 and est.LoggingTensorHook().every_n_iter == 100

    class Exploding(est.SessionRunHook):
        def after_run(self, run_context, run_values):
            run_context.estimator._last_loss_t = torch.tensor(float("nan"))

    with pytest.raises(RuntimeError, match="NaN loss"):
        make("b").train(input_fn, hooks=[Exploding(), est.NanTensorHook("loss")], max_steps=50)
    soft = make("c")
    soft.train(input_fn, hooks=[Exploding(), est.NanTensorHook("loss", fail_on_nan_loss=False)], max_steps=50)
    assert soft.get_global_step() == 1                                          # stopped at the first NaN


def test_ps_role_serves_the_plane_the_application_agreed_on(monkeypatch):
    """A ps task with a GPU still serves shared memory when the application runs the shm plane (TFY_PS_PLANE=shm)."""
    import torch
    from tf_yarn_b200.estimator import ps, ps_hbm, training
    monkeypatch.setenv("TF_CONFIG", '{"cluster": {"chief": ["a:1"], "ps": ["d:4"]}, "task": {"type": "ps", "index": 0}}')
    monkeypatch.setenv("TFY_GPU_IDS", "0")
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    served = []
    monkeypatch.setattr(ps, "serve", lambda cluster: served.append("shm"))
    monkeypatch.setattr(ps_hbm, "serve", lambda cluster: served.append("hbm"))
    monkeypatch.setenv("TFY_PS_PLANE", "shm")
    training.train_and_evaluate(None, None, None)
    monkeypatch.delenv("TFY_PS_PLANE")
    training.train_and_evaluate(None, None, None)
    assert served == ["shm", "hbm", "shm"]          # (after the hbm stub returns, the code falls through to ps.serve)
