"""Multi-process integration runs through the real launcher (CPU, gloo / shared memory).

What the reference can only exercise by hand on a YARN cluster: a full application with
several task processes, the KV rendezvous, lifecycle events, Metrics, failure propagation and retries.
"""
import os

import pytest

from tf_yarn_b200 import RunFailed, TaskSpec
from tf_yarn_b200.topologies import ContainerKey

pytestmark = pytest.mark.slow

# the task processes cannot import this test module: ship its functions by value
import sys  # noqa: E402

import cloudpickle  # noqa: E402

cloudpickle.register_pickle_by_value(sys.modules[__name__])


def _allreduce_fn(local_rank):
    import os
    import torch
    import torch.distributed as dist
    from tf_yarn_b200.distributed.task import get_task
    p = get_task(local_rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = p.master_address, str(p.master_port)
    dist.init_process_group("gloo", rank=p.task_id, world_size=p.world_size)
    t = torch.ones(1) * p.task_id
    dist.all_reduce(t)
    assert t.item() == sum(range(p.world_size))
    dist.destroy_process_group()


def test_distributed_run_two_instances_two_procs():
    from tf_yarn_b200.distributed import run_on_yarn
    metrics = run_on_yarn(_allreduce_fn, {"worker": TaskSpec("1 GiB", 2, instances=2, nb_proc_per_worker=2)})
    assert set(metrics.container_duration) == {ContainerKey("worker", 0), ContainerKey("worker", 1)}
    assert metrics.total_training_duration is not None and metrics.total_eval_duration is None


def _failing_fn(local_rank):
    raise RuntimeError("user code exploded")


def test_failure_raises_run_failed_with_the_traceback():
    from tf_yarn_b200.distributed import run_on_yarn
    with pytest.raises(RunFailed) as err:
        run_on_yarn(_failing_fn, {"worker": TaskSpec("1 GiB", 1)})
    assert "user code exploded" in str(err.value)
    assert "worker:0: FAILED" in str(err.value)


def _flaky_fn_factory(marker):
    def fn(local_rank):
        import os
        n_try = int(os.environ["TF_YARN_N_TRY"])
        with open(marker, "a") as f:
            f.write(f"{n_try}\n")
        if n_try == 0:
            raise RuntimeError("first attempt fails")
    return fn


def test_whole_application_is_retried(tmp_path):
    from tf_yarn_b200.distributed import run_on_yarn
    marker = str(tmp_path / "attempts")
    metrics = run_on_yarn(_flaky_fn_factory(marker), {"worker": TaskSpec("1 GiB", 1)}, nb_retries=1)
    assert open(marker).read().split() == ["0", "1"]
    assert metrics is not None


def test_files_env_and_pre_script_hook_reach_the_tasks(tmp_path):
    from tf_yarn_b200.distributed import run_on_yarn
    mod = tmp_path / "shipped_module.py"
    mod.write_text("VALUE = 41\n")
    out = str(tmp_path / "out.txt")

    def fn(local_rank):
        import os
        import shipped_module          # importable: the task workdir is on PYTHONPATH
        with open(os.environ["OUT_FILE"], "w") as f:
            f.write(f"{shipped_module.VALUE + 1} {os.environ['FROM_HOOK']}")
    run_on_yarn(fn, {"worker": TaskSpec("1 GiB", 1)}, files={"shipped_module.py": str(mod)},
                env={"OUT_FILE": out}, pre_script_hook="export FROM_HOOK=hooked")
    assert open(out).read() == "42 hooked"


def _keras_experiment(model_dir):
    def experiment_fn():
        import os
        import torch
        from tf_yarn_b200 import data, hvd, keras
        from tf_yarn_b200.tensorflow import KerasExperiment
        g = torch.Generator().manual_seed(int(os.environ.get("HOROVOD_RANK", "0")))
        x = torch.randn(256, 8, generator=g)
        y = (x[:, 0] > 0).long()
        model = keras.Sequential([keras.layers.Dense(16, activation="relu", input_shape=(8,)),
                                  keras.layers.Dense(2, activation="softmax")])
        opt = hvd.DistributedOptimizer(keras.optimizers.Adadelta(1.0 * int(os.environ.get("HOROVOD_SIZE", "1"))))
        model.compile(loss="sparse_categorical_crossentropy", optimizer=opt, metrics=["accuracy"])
        return KerasExperiment(
            model=model, model_dir=model_dir,
            train_params={"steps_per_epoch": 4, "epochs": 2, "verbose": 0,
                          "callbacks": [keras.callbacks.ModelCheckpoint(model_dir + "/checkpoint-{epoch}"),
                                        hvd.keras.callbacks.BroadcastGlobalVariablesCallback(0)]},
            input_data_fn=lambda: data.Dataset.from_tensor_slices((x, y)).batch(32).repeat(),
            target_data_fn=None,
            validation_data_fn=lambda: data.Dataset.from_tensor_slices((x, y)).batch(64))
    return experiment_fn


def test_keras_allreduce_chief_worker_evaluator_tensorboard(tmp_path, caplog):
    """BASELINE config 1: KerasExperiment, 1 chief + 1 worker on CPU/gloo, evaluator + TensorBoard side tasks."""
    import logging
    from tf_yarn_b200.tensorflow import run_on_yarn
    model_dir = str(tmp_path / "model")
    with caplog.at_level(logging.INFO):
        metrics = run_on_yarn(
            _keras_experiment(model_dir),
            {"chief": TaskSpec("2 GiB", 2), "worker": TaskSpec("2 GiB", 2), "evaluator": TaskSpec("2 GiB", 1),
             "tensorboard": TaskSpec("1 GiB", 1, tb_termination_timeout_seconds=1, tb_model_dir=model_dir)},
            env={"TFY_KERAS_EVAL_POLL_SECS": "0.5"},
            custom_task_module="tf_yarn_b200.tensorflow.tasks.gloo_allred_task")
    assert sorted(f for f in os.listdir(model_dir) if f.startswith("checkpoint-")) == ["checkpoint-1", "checkpoint-2"]
    assert os.path.isdir(os.path.join(model_dir, "eval"))        # the evaluator picked the checkpoints up
    assert metrics.total_training_duration is not None and metrics.total_eval_duration is not None
    assert any("Tensorboard listening on http://" in r.getMessage() for r in caplog.records)
    from tf_yarn_b200.estimator import summary
    assert sorted(set(summary.read_scalars(os.path.join(model_dir, "eval"))["step"])) == [1, 2]


def _ps_experiment(model_dir):
    def experiment_fn():
        import torch
        from tf_yarn_b200 import data, keras
        from tf_yarn_b200 import estimator as est
        from tf_yarn_b200.tensorflow import Experiment
        fc = est.feature_column
        g = torch.Generator().manual_seed(0)
        x = torch.randn(1024, 6, generator=g)
        y = (x @ torch.randn(6, 3, generator=g)).argmax(1)
        train = lambda: data.Dataset.from_tensor_slices(({"x": x}, y)).shuffle(500).batch(64).repeat()  # noqa: E731
        evalf = lambda: data.Dataset.from_tensor_slices(({"x": x}, y)).batch(256)  # noqa: E731
        e = est.DNNClassifier([16], [fc.numeric_column("x", shape=(6,))], model_dir=model_dir, n_classes=3,
                              optimizer=lambda: keras.optimizers.Adagrad(0.1),
                              config=est.RunConfig(save_checkpoints_steps=100, log_step_count_steps=None))
        return Experiment(e, est.TrainSpec(train, max_steps=200),
                          est.EvalSpec(evalf, steps=None, start_delay_secs=0, throttle_secs=0))
    return experiment_fn


def test_parameter_server_strategy_end_to_end(tmp_path):
    """chief + 2 workers + 2 ps + evaluator: async pull/push through the ps shards, stop barrier lets ps exit."""
    from tf_yarn_b200.estimator import summary
    from tf_yarn_b200.tensorflow import run_on_yarn
    import glob
    from tf_yarn_b200.estimator import ps as ps_cpu
    shards_before = set(glob.glob(os.path.join(ps_cpu._shm_dir(), "tfy_ps_*")))
    model_dir = str(tmp_path / "model")
    metrics = run_on_yarn(_ps_experiment(model_dir),
                          {"chief": TaskSpec("1 GiB", 1), "worker": TaskSpec("1 GiB", 1, instances=2),
                           "ps": TaskSpec("1 GiB", 1, instances=2), "evaluator": TaskSpec("1 GiB", 1)})
    assert metrics.container_duration[ContainerKey("ps", 1)] is not None
    # the ps tasks unlink their shared-memory shards on the way out (normal exit or the launcher's SIGTERM)
    assert set(glob.glob(os.path.join(ps_cpu._shm_dir(), "tfy_ps_*"))) <= shards_before
    sc = summary.read_scalars(os.path.join(model_dir, "eval"))
    acc = [v for n, v in zip(sc["name"], sc["value"]) if n == "accuracy"]
    # the first evaluated checkpoint may already be a late one (asynchronous evaluator): require a trained model,
    # not a strict improvement between two evaluations
    assert acc[-1] > 0.8 and acc[-1] >= acc[0] - 0.02, acc
    assert max(sc["step"]) >= 200


def test_keras_experiment_is_rejected_by_the_ps_task(tmp_path):
    from tf_yarn_b200.tensorflow import run_on_yarn
    with pytest.raises(RunFailed) as err:
        run_on_yarn(_keras_experiment(str(tmp_path / "m")), {"chief": TaskSpec("1 GiB", 1)})
    assert "KerasExperiment using parameter strategy is unsupported" in str(err.value)


_ALL_EXAMPLES = os.environ.get("TFY_TEST_ALL_EXAMPLES") == "1"      # the other two take a minute more


@pytest.mark.parametrize("example", ["mnist_cnn_allreduce", "bert_allreduce"] +
                         (["wide_deep_ps", "resnet50_ddp"] if _ALL_EXAMPLES else []))
def test_baseline_config_examples_run_end_to_end(example):
    """The four BASELINE.json configurations as a user would launch them (toy sizes on CPU): all-reduce Keras with an
    evaluator fed by `validation_data_fn=lambda: (x, y)`, the parameter-server Estimator, the PyTorch DDP experiment,
    and BERT (dict inputs / outputs) with evaluator + TensorBoard side tasks."""
    import subprocess
    import sys
    env = dict(os.environ, EXAMPLE_SMALL="1", EXAMPLE_EPOCHS="1", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, "-m", f"tf_yarn_b200.examples.baseline.{example}"], env=env,
                         capture_output=True, text=True, timeout=600,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert res.returncode == 0, (res.stdout + res.stderr)[-4000:]
    assert "Metrics(" in res.stdout
