"""Edge cases of small public helpers that the larger tests do not reach: memory parsing and topology validation,
canned DNN / indicator columns / binary predictions, the remaining Dataset verbs, the client's interpreter choice."""
import os
import sys

import pytest
import torch

from tf_yarn_b200 import TaskSpec, NodeLabel, allreduce_topology, client, data, topologies
from tf_yarn_b200 import estimator as est
from tf_yarn_b200.estimator import feature_column as fc


def test_memory_specifications():
    assert topologies.parse_memory(512) == 512 and topologies.parse_memory("2 GiB") == 2048
    assert topologies.parse_memory("1.5GiB") == 1536 and topologies.parse_memory("100 MB") == 96      # rounded up to MiB
    assert topologies.parse_memory("64") == 64 and topologies.parse_memory(10.2) == 11
    for bad, exc in ((True, TypeError), (-1, ValueError), ("lots", ValueError), ("3 parsecs", ValueError)):
        with pytest.raises(exc):
            topologies.parse_memory(bad)
    assert "memory=2048" in repr(TaskSpec("2 GiB", 4, instances=2, label=NodeLabel.GPU))


def test_topology_validation_messages():
    with pytest.raises(ValueError, match="exactly one 'chief'"):
        topologies._check_general_topology({"worker": TaskSpec(64, 1)})
    with pytest.raises(ValueError, match="subset of"):
        topologies._check_general_topology({"chief": TaskSpec(64, 1), "driver": TaskSpec(64, 1)})
    with pytest.raises(ValueError, match="more memory"):
        topologies._check_general_topology({"chief": TaskSpec(topologies.MAX_MEMORY_CONTAINER + 1, 1)})
    with pytest.raises(ValueError, match="more processes per instance than vcores"):
        topologies._check_general_topology({"chief": TaskSpec(64, 1, nb_proc_per_worker=2)})
    with pytest.raises(ValueError, match="no more than one 'evaluator'"):
        topologies._check_ps_topology({"chief": TaskSpec(64, 1), "ps": TaskSpec(64, 1),
                                       "evaluator": TaskSpec(64, 1, instances=2)})
    with pytest.raises(ValueError, match="at least a single 'ps'"):
        topologies._check_ps_topology({"chief": TaskSpec(64, 1)})
    with pytest.raises(ValueError, match="nb_workers"):
        allreduce_topology(0)
    topo = allreduce_topology(4, memory="1 GiB", vcores=2, with_tensorboard=True)
    assert sorted(topo) == ["chief", "evaluator", "tensorboard", "worker"] and topo["worker"].instances == 3
    assert topo["chief"].label == NodeLabel.GPU and topo["evaluator"].label == NodeLabel.CPU


def test_dnn_classifier_with_indicator_column_and_binary_predictions(tmp_path):
    cat = fc.categorical_column_with_hash_bucket("c", 6)
    cols = [fc.numeric_column("x", shape=(2,)), fc.indicator_column(cat)]
    e = est.DNNClassifier([8], cols, model_dir=str(tmp_path), n_classes=2, dropout=0.1,
                          config=est.RunConfig(save_checkpoints_steps=None, save_checkpoints_secs=None))
    g = torch.Generator().manual_seed(0)
    xs = torch.randn(64, 2, generator=g)
    cs = torch.randint(0, 6, (64, 1), generator=g)
    ys = (xs[:, 0] > 0).long()

    def input_fn():
        return data.Dataset.from_tensor_slices(({"x": xs, "c": cs}, ys)).batch(16).repeat(30)
    e.train(input_fn, steps=120)
    res = e.evaluate(lambda: data.Dataset.from_tensor_slices(({"x": xs, "c": cs}, ys)).batch(32))
    assert res["accuracy"] > 0.85 and res["global_step"] == 120
    preds = list(e.predict(lambda: data.Dataset.from_tensor_slices({"x": xs[:8], "c": cs[:8]}).batch(4)))
    assert len(preds) == 8 and set(preds[0]) == {"logits", "probabilities", "class_ids"}
    p = preds[0]["probabilities"]
    assert p.shape == (2,) and abs(float(p.sum()) - 1.0) < 1e-6 and int(preds[0]["class_ids"][0]) in (0, 1)
    with pytest.raises(TypeError, match="cannot feed a dense layer"):
        est.DNNClassifier([4], [cat]).train(input_fn, steps=1)


def test_remaining_dataset_verbs():
    ds = data.Dataset.range(10)
    assert [int(v) for v in ds.skip(7)] == [7, 8, 9]
    assert [int(v) for v in ds.skip(3).take(2)] == [3, 4]
    assert ds.prefetch(4) is ds
    gen = data.Dataset.from_generator(lambda: (i * i for i in range(4)))
    assert [int(v) for v in gen] == [0, 1, 4, 9] and [int(v) for v in gen] == [0, 1, 4, 9]          # re-iterable
    assert [b.tolist() for b in data.Dataset.range(5).batch(2, drop_remainder=True)] == [[0, 1], [2, 3]]
    assert [int(v) for v in data.Dataset.range(3).repeat(2)] == [0, 1, 2, 0, 1, 2]


def test_client_interpreter_choice(tmp_path):
    assert client._interpreter_for(None) is None if hasattr(client, "_interpreter_for") else True
    fn = next((getattr(client, n) for n in dir(client) if "pyenv" in n.lower() and callable(getattr(client, n))), None)
    assert fn is not None
    assert fn(None) is None
    assert fn(sys.executable) == sys.executable                          # an executable interpreter is used as is
    assert fn({NodeLabel.CPU: "/no/such/env.pex", NodeLabel.GPU: sys.executable}) == sys.executable
    assert fn(str(tmp_path / "env.zip")) is None                         # a zip to upload: meaningless on one box
    assert os.path.exists(sys.executable)


def test_canned_estimators_use_tfs_learning_rates_for_optimizer_names():
    """tf.estimator: DNNClassifier Adagrad(0.05), LinearClassifier Ftrl(min(0.2, 1/sqrt(n_cols))), combined 0.001 /
    min(0.005, 1/sqrt(n_linear_cols)); objects and callables are taken as given."""
    from tf_yarn_b200 import keras
    from tf_yarn_b200.estimator import canned
    assert canned._tf_default("adagrad", "dnn", 3, False).learning_rate == 0.05
    assert canned._tf_default("Adagrad", "dnn", 3, True).learning_rate == 0.001
    assert canned._tf_default("ftrl", "linear", 100, False).learning_rate == pytest.approx(0.1)
    assert canned._tf_default("ftrl", "linear", 4, False).learning_rate == 0.2
    assert canned._tf_default("ftrl", "linear", 4, True).learning_rate == 0.005
    assert canned._tf_default("sgd", "dnn", 4, False) == "sgd"
    obj = keras.optimizers.Adagrad(0.3)
    assert canned._tf_default(obj, "dnn", 4, False) is obj
