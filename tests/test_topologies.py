import pytest

from tf_yarn_b200 import topologies
from tf_yarn_b200.topologies import (ContainerKey, ContainerTask, NodeLabel, TaskSpec, parse_memory,
                                     ps_strategy_topology, single_server_topology)

sock_addrs = {"chief": ["addr1:port1"], "evaluator": ["addr5:port5"], "ps": ["addr6:port6"],
              "worker": ["addr7:port7", "addr8:port8"]}


def test_single_server_topology():
    topo = single_server_topology()
    assert set(topo) == {"chief", "evaluator", "tensorboard"}
    assert all(s.instances == 1 for s in topo.values())


def test_ps_strategy_topology():
    topo = ps_strategy_topology(nb_workers=4, nb_ps=2)
    assert topo["worker"].instances == 4 and topo["ps"].instances == 2
    assert topo["chief"].instances == 1


def test_memory_and_vcores_caps():
    with pytest.raises(ValueError):
        single_server_topology(memory=topologies.MAX_MEMORY_CONTAINER + 1)
    with pytest.raises(ValueError):
        single_server_topology(vcores=topologies.MAX_VCORES_CONTAINER + 1)
    with pytest.raises(ValueError):
        ps_strategy_topology(nb_ps=2, memory=topologies.MAX_MEMORY_CONTAINER + 1)


def test_ps_topology_needs_a_ps():
    with pytest.raises(ValueError):
        ps_strategy_topology(nb_ps=0)


def test_unknown_task_type_rejected():
    with pytest.raises(ValueError):
        topologies._check_general_topology({"chief": TaskSpec(1, 1), "driver": TaskSpec(1, 1)})


def test_more_procs_than_vcores_rejected():
    with pytest.raises(ValueError):
        topologies._check_general_topology({"chief": TaskSpec(1024, 1, nb_proc_per_worker=2)})


@pytest.mark.parametrize("value,expected", [(2048, 2048), ("2 GiB", 2048), ("512 MiB", 512), ("1GiB", 1024),
                                            ("1 GB", 954), ("100", 100)])
def test_parse_memory(value, expected):
    assert parse_memory(value) == expected


def test_taskspec_properties():
    spec = TaskSpec("2 GiB", 4, instances=3, nb_proc_per_worker=2, label=NodeLabel.GPU)
    assert (spec.memory, spec.vcores, spec.instances, spec.nb_proc_per_worker) == (2048, 4, 3, 2)
    spec.memory = "1 GiB"
    spec.vcores = 8
    assert (spec.memory, spec.vcores) == (1024, 8)
    assert NodeLabel.CPU.value == "" and NodeLabel.GPU.value == "gpu"


def test_container_key_round_trip():
    key = ContainerKey("worker", 3)
    assert key.to_kv_str() == "worker:3"
    assert ContainerKey.from_kv_str("worker:3") == key
    assert ContainerKey.from_kv_str("garbage") is None
    task = ContainerTask("ps", 1, 2)
    assert task.to_container_key() == ContainerKey("ps", 1) and task.to_kv_str() == "ps:1:2"


def test_allreduce_topology():
    topo = topologies.allreduce_topology(8)
    assert topo["chief"].instances == 1 and topo["worker"].instances == 7
    assert topo["chief"].label == NodeLabel.GPU
