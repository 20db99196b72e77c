"""Client-side driver: cluster spec, retries, event aggregation/reporting, safe experiment fn."""
import json
from datetime import timedelta
from unittest import mock

import pytest

from tf_yarn_b200 import client, constants
from tf_yarn_b200.client import ContainerLogStatus, RunFailed, _handle_events, get_safe_experiment_fn, run_on_yarn
from tf_yarn_b200.topologies import ContainerKey, NodeLabel, TaskSpec

from fakes import FakeClient


def test_setup_cluster_spec_excludes_side_tasks():
    app = FakeClient()
    client._setup_cluster_spec([("chief", 1, 1), ("worker", 2, 4), ("ps", 1, 1), ("evaluator", 1, 1),
                                ("tensorboard", 1, 1)], app)
    assert json.loads(app.kv[constants.KV_CLUSTER_INSTANCES]) == [["chief", 1, 1], ["worker", 2, 4], ["ps", 1, 1]]


def test_setup_to_use_cuda_archive_needs_ld_library_path():
    hook = client._setup_to_use_cuda_archive({}, "echo hi", "hdfs://x/cuda.tar")
    assert hook == "echo hi"
    hook = client._setup_to_use_cuda_archive({"LD_LIBRARY_PATH": "cuda/lib64"}, "echo hi", "hdfs://x/cuda.tar")
    assert hook == "echo hi"      # the box's toolkit is used; the hook is left untouched


def test_nb_retries_must_be_non_negative():
    with pytest.raises(ValueError):
        run_on_yarn(lambda: None, {"chief": TaskSpec(1024, 1)}, nb_retries=-1)


@pytest.mark.parametrize("nb_retries,nb_failures", [(0, 0), (1, 0), (1, 1), (2, 2), (0, 1), (3, 1)])
def test_retry_run_on_yarn(nb_retries, nb_failures):
    state = {"calls": 0}

    def maybe_fail(*args, **kwargs):
        state["calls"] += 1
        if state["calls"] <= nb_failures:
            raise RuntimeError("attempt failed")
        return "metrics"

    fake_cluster = mock.MagicMock()
    with mock.patch("tf_yarn_b200.client._setup_skein_cluster", return_value=fake_cluster) as setup, \
            mock.patch("tf_yarn_b200.client._run_on_cluster", side_effect=maybe_fail) as run:
        result = None
        try:
            result = run_on_yarn(lambda: None, {"chief": TaskSpec(1024, 1)}, nb_retries=nb_retries)
        except RuntimeError:
            pass
        expected_calls = min(nb_retries, nb_failures) + 1
        assert run.call_count == expected_calls
        assert setup.call_count == expected_calls
        assert (result == "metrics") == (nb_failures <= nb_retries)
        # the attempt index is handed to every new cluster
        assert [c.kwargs["n_try"] for c in setup.call_args_list] == list(range(expected_calls))


def test_app_is_shut_down_when_the_experiment_cannot_be_pickled():
    app = FakeClient()
    fake_cluster = mock.MagicMock()
    fake_cluster.app = app
    fake_cluster.tasks = [("chief", 1, 1)]
    with mock.patch("tf_yarn_b200.client._setup_skein_cluster", return_value=fake_cluster), \
            mock.patch("tf_yarn_b200.client.cloudpickle.dumps", side_effect=TypeError("not picklable")):
        with pytest.raises(TypeError):
            run_on_yarn(lambda: None, {"chief": TaskSpec(1024, 1)})
    assert app.shutdown_status == "failed"


def test_get_safe_experiment_fn():
    fn = get_safe_experiment_fn("os.path.join", "a", "b")
    assert fn() == "a/b"
    with mock.patch("importlib.import_module") as imp:
        imp.return_value.my_fn.return_value = "exp"
        assert get_safe_experiment_fn("pkg.mod.my_fn", 1)() == "exp"
        imp.assert_called_with("pkg.mod")
        imp.return_value.my_fn.assert_called_with(1)


def test_container_log_status():
    status = ContainerLogStatus(
        {ContainerKey("chief", 0): "http://host:8042/node/containerlogs/container_e1_01_000002/myuser",
         ContainerKey("evaluator", 0): "/tmp/app/logs/container_evaluator_0/task.log"},
        {ContainerKey("chief", 0): "SUCCEEDED", ContainerKey("evaluator", 0): "FAILED"})
    containers = status.by_container_id()
    assert containers["container_e1_01_000002"] == (ContainerKey("chief", 0), "SUCCEEDED")
    assert containers["container_evaluator_0"] == (ContainerKey("evaluator", 0), "FAILED")
    assert ContainerLogStatus({ContainerKey("chief", 0): "x"}, {}).by_container_id() == {}


def test_handle_events_statuses_and_durations():
    events = {
        ContainerKey("chief", 0): {"init": "h:1", "start": "", "stop": "", "logs": "/l/container_chief_0/task.log",
                                   "container_start_time": "100.0", "container_stop_time": "160.0",
                                   "train_eval_start_time": "110.0", "train_eval_stop_time": "150.0"},
        ContainerKey("worker", 0): {"init": "h:2", "start": "", "stop": "",
                                    "train_eval_start_time": "105.0", "train_eval_stop_time": "155.0"},
        ContainerKey("evaluator", 0): {"init": "h:3", "stop": "Traceback: boom",
                                       "train_eval_start_time": "120.0", "train_eval_stop_time": "130.0"},
        ContainerKey("ps", 0): {"init": "h:4"},
        ContainerKey("tensorboard", 0): {},
    }
    text, metrics, statuses = _handle_events(events, 0)
    st = statuses.container_status
    assert st[ContainerKey("chief", 0)] == "SUCCEEDED"
    assert st[ContainerKey("evaluator", 0)] == "FAILED"
    assert st[ContainerKey("ps", 0)] == "KILLED"
    assert st[ContainerKey("tensorboard", 0)] == "REQUESTED"
    assert metrics.container_duration[ContainerKey("chief", 0)] == timedelta(seconds=60)
    assert metrics.train_eval_time_per_node[ContainerKey("worker", 0)] == timedelta(seconds=50)
    # max(stop) - min(start) over chief + workers
    assert metrics.total_training_duration == timedelta(seconds=50)
    # the evaluator failed => no valid evaluation time
    assert metrics.total_eval_duration is None
    assert "Exception in task evaluator:0" in text and "boom" in text


def test_handle_events_training_time_invalid_when_a_trainer_did_not_report():
    events = {
        ContainerKey("chief", 0): {"stop": "", "train_eval_start_time": "1.0", "train_eval_stop_time": "2.0"},
        ContainerKey("worker", 0): {"stop": ""},
    }
    _, metrics, _ = _handle_events(events, 0)
    assert metrics.total_training_duration is None


def test_aggregate_events_collects_task_stages():
    app = FakeClient()
    events = {ContainerKey("worker", 0): {}, ContainerKey("worker", 1): {}}
    import threading
    stop = threading.Event()
    t = threading.Thread(target=client._aggregate_events, args=(app.kv, events, stop))
    t.start()
    app.kv["worker:0/init"] = b"1.2.3.4:5"
    app.kv["worker:1/stop"] = b""
    app.kv["cluster_instances"] = b"[]"       # not a task event
    app.kv["ghost:7/init"] = b"x"             # unknown task: ignored
    import time
    time.sleep(0.5)
    stop.set()
    t.join(5)
    assert events == {ContainerKey("worker", 0): {"init": "1.2.3.4:5"}, ContainerKey("worker", 1): {"stop": ""}}


def test_run_failed_is_an_exception():
    assert issubclass(RunFailed, Exception)


def test_services_carry_the_task_environment_and_scripts(monkeypatch):
    """What each role is started with: retry index, PYTHONPATH for shipped files, TensorBoard knobs from the TaskSpec,
    the pre-script hook ahead of the task command, the custom task module for trainers only
    (reference: tf_yarn/client.py:108-151, _env.py:6-24)."""
    captured = {}

    class StubApp:
        def __init__(self):
            from tf_yarn_b200.kv import InMemoryKV
            self.kv = InMemoryKV()

    class StubClient:
        def submit_and_connect(self, spec):
            captured["spec"] = spec
            return StubApp()

    specs = {"chief": TaskSpec("1 GiB", 2, label=NodeLabel.GPU), "worker": TaskSpec("1 GiB", 2, instances=3),
             "tensorboard": TaskSpec("512 MiB", 1, tb_model_dir="/models/m", tb_extra_args="--reload_interval 5",
                                     tb_termination_timeout_seconds=7)}
    cluster = client._setup_local_cluster(
        specs, custom_task_module="my_pkg.my_task", local_client=StubClient(), files={"my_pkg": "/src/my_pkg"},
        env={"LD_LIBRARY_PATH": "/usr/lib", "FOO": "bar"}, n_try=2, pre_script_hook="source env.sh",
        cuda_runtime_hdfs_path="viewfs://root/cuda.tar.gz", name="job", queue="ml")
    cluster.event_listener.stop_event.set()
    app_spec = captured["spec"]
    assert app_spec.name == "job" and app_spec.queue == "ml" and set(app_spec.services) == set(specs)
    chief, worker, tb = (app_spec.services[k] for k in ("chief", "worker", "tensorboard"))
    assert chief.env["TF_YARN_N_TRY"] == chief.env["TFY_N_TRY"] == "2" and chief.env["FOO"] == "bar"
    assert chief.env["PYTHONPATH"].startswith(".:") and chief.files == {"my_pkg": "/src/my_pkg"}
    assert (worker.instances, worker.nb_proc, worker.vcores, worker.memory) == (3, 1, 2, 1024)
    assert chief.label == NodeLabel.GPU and worker.label == NodeLabel.CPU
    assert chief.script.splitlines()[:2] == ["set -e", "source env.sh"]
    assert chief.script.splitlines()[-1].startswith("exec ") and "my_pkg.my_task" in chief.script
    assert "my_pkg.my_task" not in tb.script and "_tensorboard_task" in tb.script
    assert tb.env["TB_MODEL_DIR"] == "/models/m" and tb.env["TB_EXTRA_ARGS"] == "--reload_interval 5"
    assert tb.env["TB_TERMINATION_TIMEOUT_SECONDS"] == "7" and "TB_MODEL_DIR" not in chief.env
    assert cluster.tasks == [("chief", 1, 1), ("worker", 3, 1), ("tensorboard", 1, 1)]
    assert client._default_acls_all_access() == {"enable": True, "ui_users": ["*"], "view_users": ["*"]}


def test_one_parameter_server_plane_per_application():
    """The peer-HBM plane needs every trainer AND ps on a GPU.  The reference's default ps_strategy_topology() labels
    nothing: on a GPU box the trainers would otherwise pick the HBM plane while the ps serves shared memory (a hang)."""
    from tf_yarn_b200 import topologies

    def plane(specs):
        captured = {}

        class StubApp:
            def __init__(self):
                from tf_yarn_b200.kv import InMemoryKV
                self.kv = InMemoryKV()

        class StubClient:
            def submit_and_connect(self, spec):
                captured["spec"] = spec
                return StubApp()
        cluster = client._setup_local_cluster(specs, local_client=StubClient())
        cluster.event_listener.stop_event.set()
        return {k: v.env.get("TFY_PS_PLANE") for k, v in captured["spec"].services.items()}

    gpu = lambda n=1: TaskSpec("1 GiB", 1, instances=n, label=NodeLabel.GPU)   # noqa: E731
    cpu = lambda n=1: TaskSpec("1 GiB", 1, instances=n)                         # noqa: E731
    assert set(plane(topologies.ps_strategy_topology(2, 1, "1 GiB", 1)).values()) == {"shm"}
    assert set(plane({"chief": gpu(), "worker": gpu(2), "ps": cpu()}).values()) == {"shm"}
    assert set(plane({"chief": gpu(), "worker": gpu(2), "ps": gpu(), "evaluator": cpu()}).values()) == {None}
    assert set(plane({"chief": gpu(), "worker": cpu()}).values()) == {None}                    # no ps: not a PS job
