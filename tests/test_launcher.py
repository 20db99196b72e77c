"""Local launcher (skein/YARN stand-in): process-group lifecycle, environment, CPU pinning, final status."""
import os
import time

import pytest

from tf_yarn_b200.launcher import local
from tf_yarn_b200.topologies import NodeLabel


def _svc(script, instances=1, vcores=1, label=NodeLabel.CPU, env=None):
    return local.ServiceSpec(script=script, instances=instances, nb_proc=1, label=label, memory=64, vcores=vcores,
                             env=env or {}, files={})


def _alive(pid: int) -> bool:
    try:
        os.kill(pid, 0)
    except ProcessLookupError:
        return False
    except PermissionError:
        return True
    # a zombie still answers kill(0): look at its state
    try:
        with open(f"/proc/{pid}/stat") as f:
            return f.read().split(")")[-1].split()[0] != "Z"
    except OSError:
        return False


def _wait(cond, timeout=20.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if cond():
            return True
        time.sleep(0.1)
    return False


def test_shutdown_kills_the_whole_process_group_of_every_task(tmp_path):
    """A task's children (data-loader workers, spawned trainers) must not outlive the application."""
    script = "sleep 300 & echo child=$! > pids.txt; echo self=$$ >> pids.txt; wait"
    app = local.LocalClient(str(tmp_path)).submit_and_connect(
        local.ApplicationSpec({"worker": _svc(script, instances=2)}, name="orphans"))
    assert _wait(lambda: all(os.path.exists(os.path.join(p.workdir, "pids.txt")) and
                             len(open(os.path.join(p.workdir, "pids.txt")).read().split()) == 2 for p in app.processes))
    pids = []
    for p in app.processes:
        for line in open(os.path.join(p.workdir, "pids.txt")).read().split():
            pids.append(int(line.split("=")[1]))
    assert len(pids) == 4 and all(_alive(pid) for pid in pids)
    assert app.report().state == "running"
    app.shutdown(local.FinalStatus.KILLED)
    assert _wait(lambda: not any(_alive(pid) for pid in pids)), [pid for pid in pids if _alive(pid)]
    rep = app.report()
    assert rep.state == "killed" and rep.final_status == "killed" and rep.finish_time >= rep.start_time
    app.cleanup()


def test_a_failing_task_fails_the_application_and_stops_the_others(tmp_path):
    app = local.LocalClient(str(tmp_path)).submit_and_connect(local.ApplicationSpec(
        {"chief": _svc("sleep 0.5; echo boom >&2; exit 3"), "worker": _svc("sleep 300", instances=2)}, name="fail"))
    assert _wait(lambda: app.report().state == "failed")
    assert _wait(lambda: all(p.returncode is not None for p in app.processes))
    logs = app.logs()
    assert "boom" in logs["container_chief_0"] and set(logs) == {"container_chief_0", "container_worker_0",
                                                                 "container_worker_1"}
    app.close()


def test_success_needs_every_task_to_exit_zero_and_tasks_see_their_environment(tmp_path):
    script = ('echo "key=$TFY_TASK_KEY mem=$TFY_MEMORY_MB vcores=$TFY_VCORES cpus=${TFY_CPUS:-none} '
              'gpus=[$TFY_GPU_IDS] extra=$EXTRA kv=$TFY_KV_ADDR"')
    app = local.LocalClient(str(tmp_path)).submit_and_connect(local.ApplicationSpec(
        {"chief": _svc(script, vcores=1, env={"EXTRA": "42"}), "evaluator": _svc(script, vcores=1)}, name="ok"))
    assert _wait(lambda: app.report().state == "finished")
    assert app.report().final_status == "succeeded"
    out = app.logs()["container_chief_0"]
    assert "key=chief:0 mem=64 vcores=1" in out and "extra=42" in out and "kv=127.0.0.1:" in out
    if hasattr(os, "sched_getaffinity") and len(os.sched_getaffinity(0)) >= 2:
        a = out.split("cpus=")[1].split()[0]
        b = app.logs()["container_evaluator_0"].split("cpus=")[1].split()[0]
        assert a != "none" and b != "none" and a != b                    # disjoint CPU sets
    app.close()
    app.cleanup()


def test_visible_gpus_override(monkeypatch):
    monkeypatch.setenv("TFY_VISIBLE_GPUS", "2, 5,7")
    assert local.visible_gpus() == [2, 5, 7]
    monkeypatch.setenv("TFY_VISIBLE_GPUS", "")
    assert local.visible_gpus() == []


@pytest.mark.parametrize("n_gpu_tasks,gpus,expect", [(2, [0, 1, 2, 3], [[0], [1]]), (3, [4, 5], [[4], [5], [4]])])
def test_gpu_placement_round_robin(tmp_path, monkeypatch, n_gpu_tasks, gpus, expect):
    monkeypatch.setenv("TFY_VISIBLE_GPUS", ",".join(map(str, gpus)))
    app = local.LocalClient(str(tmp_path)).submit_and_connect(local.ApplicationSpec(
        {"worker": _svc("true", instances=n_gpu_tasks, label=NodeLabel.GPU), "evaluator": _svc("true")}, name="place"))
    assert [app.placement[f"worker:{i}"] for i in range(n_gpu_tasks)] == expect and app.placement["evaluator:0"] == []
    assert _wait(lambda: app.report().state == "finished")
    app.close()


def test_cpu_labelled_tasks_do_not_see_the_gpus(tmp_path, monkeypatch):
    """No GPU label = host CPUs only (a YARN node without GPUs): the evaluator must not compute on the chief's GPU,
    and a CPU-labelled ps / trainer must not pick the GPU-only data plane.  TFY_CPU_TASKS_SEE_GPUS=1 restores access."""
    monkeypatch.setenv("TFY_VISIBLE_GPUS", "0,1")
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "0,1")
    script = 'echo "cvd=[${CUDA_VISIBLE_DEVICES-unset}] gpus=[$TFY_GPU_IDS]"'
    spec = local.ApplicationSpec({"chief": _svc(script, label=NodeLabel.GPU), "evaluator": _svc(script)}, name="vis")
    app = local.LocalClient(str(tmp_path / "a")).submit_and_connect(spec)
    assert _wait(lambda: app.report().state == "finished")
    assert "cvd=[0,1] gpus=[0]" in app.logs()["container_chief_0"]          # trainers keep every GPU visible (peer access)
    assert "cvd=[] gpus=[]" in app.logs()["container_evaluator_0"]
    app.close()
    monkeypatch.setenv("TFY_CPU_TASKS_SEE_GPUS", "1")
    app = local.LocalClient(str(tmp_path / "b")).submit_and_connect(spec)
    assert _wait(lambda: app.report().state == "finished")
    assert "cvd=[0,1]" in app.logs()["container_evaluator_0"]
    app.close()
