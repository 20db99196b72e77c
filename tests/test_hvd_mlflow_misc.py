"""hvd facade (single process), experiment tracking, tensorboard helpers, task command generation."""
import os

import pytest
import torch

from tf_yarn_b200 import _env, hvd, metrics, mlflow, tensorboard
from tf_yarn_b200.topologies import ContainerKey, ContainerTask

from fakes import FakeClient


@pytest.fixture(autouse=True)
def _fresh_hvd(monkeypatch):
    for k in list(os.environ):
        if k.startswith("HOROVOD_") or k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            monkeypatch.delenv(k)
    hvd._state.update(initialized=False, rank=0, size=1, local_rank=0, local_size=1)
    yield
    hvd._state.update(initialized=False, rank=0, size=1, local_rank=0, local_size=1)


def test_hvd_requires_init():
    with pytest.raises(ValueError):
        hvd.rank()


def test_hvd_identity_from_horovod_env(monkeypatch):
    monkeypatch.setenv("HOROVOD_RANK", "3")
    monkeypatch.setenv("HOROVOD_SIZE", "8")
    monkeypatch.setenv("HOROVOD_LOCAL_RANK", "3")
    monkeypatch.setenv("HOROVOD_LOCAL_SIZE", "8")
    hvd.init()
    assert (hvd.rank(), hvd.size(), hvd.local_rank(), hvd.local_size()) == (3, 8, 3, 8)


def test_hvd_single_process_collectives_are_identity():
    hvd.init()
    t = torch.arange(4.0)
    assert torch.equal(hvd.allreduce(t), t) and torch.equal(hvd.broadcast(t, 0), t)
    hvd.broadcast_parameters({"w": t}, 0)
    hvd.barrier()


def test_distributed_optimizer_wrappers():
    from tf_yarn_b200 import keras
    hvd.init()
    wrapped = hvd.DistributedOptimizer(keras.optimizers.Adadelta(1.0))
    desc = keras.optimizers.get(wrapped)
    assert desc.distributed and isinstance(desc, keras.optimizers.Adadelta)
    p = torch.nn.Parameter(torch.ones(2))
    opt = hvd.DistributedOptimizer(torch.optim.SGD([p], lr=0.5))
    p.grad = torch.ones(2)
    opt.step()
    assert torch.allclose(p.detach(), torch.full((2,), 0.5))
    factory = hvd.DistributedOptimizer(lambda: keras.optimizers.Adam(0.1))
    assert keras.optimizers.get(factory()).distributed
    assert isinstance(hvd.keras.callbacks.BroadcastGlobalVariablesCallback(0), keras.callbacks.Callback)


def test_local_tracker_and_metrics_artifact(tmp_path, monkeypatch):
    monkeypatch.setenv("TFY_TRACKING_DIR", str(tmp_path))
    monkeypatch.delenv("TF_YARN_USE_MLFLOW", raising=False)
    mlflow.reset()
    assert mlflow.use_mlflow()
    mlflow.log_metric("steps_per_sec_0", 12.5, step=10)
    mlflow.set_tag("tensorboard:0/url_0", "http://h:1")
    mlflow.log_param("lr", 0.1)
    m = metrics.Metrics(None, None, {ContainerKey("chief", 0): __import__("datetime").timedelta(seconds=3)}, {})
    m.log_mlflow(0)
    tracker = mlflow.backend()
    assert tracker.metric_history("steps_per_sec_0")[0]["value"] == 12.5
    assert "tensorboard_0_url_0" in tracker.list("tags")
    assert "tf_yarn_duration_stats" in tracker.list("artifacts")
    env = mlflow.task_env()
    assert env["MLFLOW_RUN_ID"] == tracker.run_id and env["TFY_TRACKING_DIR"] == str(tmp_path)
    mlflow.reset()


def test_tracking_disabled_is_a_noop(monkeypatch):
    monkeypatch.setenv("TF_YARN_USE_MLFLOW", "False")
    mlflow.reset()
    assert not mlflow.use_mlflow()
    assert mlflow.log_metric("x", 1.0) is None and mlflow.active_run_id() == ""
    mlflow.reset()


def test_format_key():
    assert mlflow.format_key("tensorboard:0/url") == "tensorboard_0_url" and mlflow.format_key("") == ""


def test_one_shot_metrics_logger_logs_once(caplog):
    app = FakeClient()
    logger = metrics.OneShotMetricsLogger(app, [("tensorboard:0/url", "Tensorboard listening on")], 0)
    logger.log()
    assert logger.events
    app.kv["tensorboard:0/url"] = b"http://h:6006"
    with caplog.at_level("INFO"):
        logger.log()
        logger.log()
    assert not logger.events
    assert sum("Tensorboard listening on http://h:6006" in r.getMessage() for r in caplog.records) == 1


def test_tensorboard_helpers(monkeypatch):
    tasks = [ContainerTask("chief", 0, 1), ContainerTask("tensorboard", 0, 1)]
    assert tensorboard.url_event_name(tasks) == "tensorboard:0/url"
    assert tensorboard.url_event_name(tasks + [ContainerTask("tensorboard", 1, 1)]) is None
    monkeypatch.delenv("TB_TERMINATION_TIMEOUT_SECONDS", raising=False)
    assert tensorboard.get_termination_timeout() == 30
    monkeypatch.setenv("TB_TERMINATION_TIMEOUT_SECONDS", "5")
    assert tensorboard.get_termination_timeout() == 5


def test_gen_task_cmd():
    assert _env.gen_task_cmd("py", "tensorboard").strip() == "py -m tf_yarn_b200.tensorflow.tasks._tensorboard_task"
    assert _env.gen_task_cmd("py", "worker").strip().endswith("_independent_workers_task")
    assert _env.gen_task_cmd("py", "ps", "my.module").strip() == "py -m my.module"
    with pytest.raises(ValueError):
        _env.gen_task_cmd("py", "driver")


def test_check_hadoop_env_entry_point_runs_the_whole_check(tmp_path, monkeypatch):
    """Same module name as the reference's environment check: file round trip + a one-task application
    that re-reads the file and posts `result` (reference: tf_yarn/bin/check_hadoop_env.py:125-172)."""
    from tf_yarn_b200.bin import check_hadoop_env
    monkeypatch.chdir(tmp_path)
    path = check_hadoop_env.write_dummy_file(str(tmp_path))
    assert check_hadoop_env.read_file(path) == check_hadoop_env.EXPECTED_CONTENT
    monkeypatch.setattr(check_hadoop_env._impl, "check_native", lambda: True)   # (built by the other tests)
    monkeypatch.setattr(check_hadoop_env._impl, "check_gpus", lambda: True)
    assert check_hadoop_env.main([]) == 0
    log = (tmp_path / check_hadoop_env.RESULT_CHECK_FILE).read_text()
    assert "remote_check: True" in log and "setup: OK" in log


def test_hvd_broadcast_variables_alias_single_rank():
    import torch
    from tf_yarn_b200 import hvd
    hvd.init()
    t = torch.arange(4.0)
    hvd.broadcast_variables([t], root_rank=0)          # one rank: nothing to exchange
    assert t.tolist() == [0.0, 1.0, 2.0, 3.0]


def test_packaging_zip_and_local_upload(tmp_path):
    """Reference: tf_yarn/packaging.py (deprecated pass-throughs): the names still work on one box."""
    import sys
    import warnings
    import zipfile
    from tf_yarn_b200 import packaging
    src = tmp_path / "proj"
    (src / "pkg").mkdir(parents=True)
    (src / "pkg" / "mod.py").write_text("X = 1\n")
    z = packaging.zip_path(str(src), include_base_name=True, tmp_dir=str(tmp_path / "out1"))
    assert sorted(zipfile.ZipFile(z).namelist()) == ["proj/pkg/mod.py"]
    z2 = packaging.zip_path(str(src), include_base_name=False, tmp_dir=str(tmp_path / "out2"))
    assert sorted(zipfile.ZipFile(z2).namelist()) == ["pkg/mod.py"]
    dst = tmp_path / "shipped" / "proj.zip"
    assert packaging.upload_zip(z, str(dst)) == str(dst) and dst.exists()
    assert packaging.upload_zip(z) == z
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        exe, name = packaging.upload_env()
    assert exe == sys.executable and w and issubclass(w[0].category, DeprecationWarning)
    assert packaging.get_editable_requirements() == {} and packaging.get_default_fs() == "file://"
