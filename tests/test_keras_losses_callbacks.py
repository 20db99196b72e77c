"""mini-Keras losses, metrics and callbacks against plain torch / hand-computed values, and the TensorBoard launcher
(reference user code: examples/native_keras_with_gloo_example.py:70-78; tf_yarn/tensorboard.py:28-49)."""
import os
import sys
import types

import pytest
import torch
import torch.nn.functional as F

from fakes import FakeClient
from tf_yarn_b200 import keras
from tf_yarn_b200.keras import losses, metrics


def test_losses_match_torch():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(6, 4, generator=g)
    y = torch.tensor([0, 1, 2, 3, 1, 0])
    probs = logits.softmax(-1)
    onehot = F.one_hot(y, 4).float()
    ref = F.cross_entropy(logits, y)
    assert torch.allclose(losses.sparse_categorical_crossentropy(y, logits, from_logits=True), ref, atol=1e-6)
    assert torch.allclose(losses.sparse_categorical_crossentropy(y, probs), ref, atol=1e-5)
    assert torch.allclose(losses.categorical_crossentropy(onehot, logits, from_logits=True), ref, atol=1e-6)
    assert torch.allclose(losses.CategoricalCrossentropy()(onehot, probs), ref, atol=1e-5)
    z = torch.randn(8, 1, generator=g)
    t = torch.tensor([0., 1, 1, 0, 1, 0, 0, 1])
    bref = F.binary_cross_entropy_with_logits(z.reshape(-1), t)
    assert torch.allclose(losses.binary_crossentropy(t, z, from_logits=True), bref, atol=1e-6)
    assert torch.allclose(losses.BinaryCrossentropy()(t, torch.sigmoid(z)), bref, atol=1e-5)
    a, b = torch.randn(5, 3, generator=g), torch.randn(5, 3, generator=g)
    assert torch.allclose(losses.mean_squared_error(a, b), F.mse_loss(b, a))
    assert torch.allclose(losses.mean_absolute_error(a, b), F.l1_loss(b, a))
    assert losses.get("mse") is not None and losses.name_of("sparse_categorical_crossentropy")
    with pytest.raises(Exception):
        losses.get("no_such_loss")
    cfg = losses.serialize(losses.SparseCategoricalCrossentropy(from_logits=True))
    assert losses.deserialize(cfg)(y, logits).item() == pytest.approx(ref.item(), abs=1e-6)
    assert losses.deserialize(losses.serialize("mse")) == "mse"        # a name round-trips as the identifier compile() accepts
    assert losses.serialize(lambda t, p: p) is None and losses.deserialize(None) is None


def test_metric_resolution_and_values():
    y = torch.tensor([0, 1, 2, 1])
    pred = torch.tensor([[.8, .1, .1], [.2, .7, .1], [.3, .4, .3], [.1, .8, .1]])
    num, den = metrics.sparse_categorical_accuracy(y, pred)
    assert float(num) / den == 0.75
    num, den = metrics.categorical_accuracy(F.one_hot(y, 3).float(), pred)
    assert float(num) / den == 0.75
    num, den = metrics.binary_accuracy(torch.tensor([1., 0, 1, 0]), torch.tensor([.9, .2, .4, .6]))
    assert float(num) / den == 0.5
    num, den = metrics.mean_absolute_error(torch.tensor([[1.], [2.]]), torch.tensor([[2.], [4.]]))
    assert float(num) / den == 1.5
    assert metrics.resolve("accuracy", "sparse_categorical_crossentropy", 3)[1] is metrics.sparse_categorical_accuracy
    assert metrics.resolve("accuracy", "categorical_crossentropy", 3)[1] is metrics.categorical_accuracy
    assert metrics.resolve("acc", "binary_crossentropy", 1)[1] is metrics.binary_accuracy
    assert metrics.resolve("mae", "mse", 1) == ("mae", metrics.mean_absolute_error)
    fn = lambda t, p: (torch.tensor(1.0), 1.0)   # noqa: E731
    assert metrics.resolve(fn, "mse", 1)[1] is fn
    with pytest.raises(ValueError):
        metrics.resolve("auc", "mse", 1)


class _StubModel:
    def __init__(self):
        self.stop_training = False
        self.saved, self.saved_weights, self.lr = [], [], 0.1

    def save(self, path):
        self.saved.append(path)

    def save_weights(self, path):
        self.saved_weights.append(path)

    def get_learning_rate(self):
        return self.lr

    def set_learning_rate(self, lr):
        self.lr = lr


def test_early_stopping_patience_follows_keras():
    cb = keras.callbacks.EarlyStopping(monitor="val_loss", patience=2)
    m = _StubModel()
    cb.set_model(m)
    stops = []
    for epoch, v in enumerate([1.0, 0.9, 0.95, 0.93, 0.5]):
        cb.on_epoch_end(epoch, {"val_loss": v})
        stops.append(m.stop_training)
    assert stops == [False, False, False, True, True]          # two epochs without improvement -> stop
    acc = keras.callbacks.EarlyStopping(monitor="val_accuracy", patience=0)
    m2 = _StubModel()
    acc.set_model(m2)
    acc.on_epoch_end(0, {"val_accuracy": 0.5})
    acc.on_epoch_end(1, {"val_accuracy": 0.6})                 # "acc" in the name: higher is better
    assert not m2.stop_training
    acc.on_epoch_end(2, {"val_accuracy": 0.6})
    assert m2.stop_training
    acc.on_epoch_end(3, {})                                    # a missing metric is ignored


def test_model_checkpoint_modes_and_lr_scheduler(tmp_path):
    m = _StubModel()
    cb = keras.callbacks.ModelCheckpoint(str(tmp_path / "w.{epoch:02d}-{val_loss:.2f}.h5"), save_best_only=True)
    cb.set_model(m)
    for epoch, v in enumerate([0.5, 0.7, 0.4]):
        cb.on_epoch_end(epoch, {"val_loss": v})
    assert [os.path.basename(p) for p in m.saved] == ["w.01-0.50.h5", "w.03-0.40.h5"]
    cb.on_epoch_end(3, {})                                     # monitored value missing: nothing saved
    assert len(m.saved) == 2
    every2 = keras.callbacks.ModelCheckpoint(str(tmp_path / "sub" / "ck-{epoch}"), save_weights_only=True, period=2)
    every2.set_model(m)
    for epoch in range(4):
        every2.on_epoch_end(epoch, {})
    assert [os.path.basename(p) for p in m.saved_weights] == ["ck-2", "ck-4"] and (tmp_path / "sub").is_dir()
    two_args = keras.callbacks.LearningRateScheduler(lambda epoch, lr: lr * 0.5)
    two_args.set_model(m)
    two_args.on_epoch_begin(0)
    assert m.lr == pytest.approx(0.05)
    one_arg = keras.callbacks.LearningRateScheduler(lambda epoch: 0.01 * (epoch + 1))
    one_arg.set_model(m)
    one_arg.on_epoch_begin(2)
    assert m.lr == pytest.approx(0.03)


def test_tensorboard_callback_writes_event_files(tmp_path):
    x = torch.randn(32, 4)
    y = (x[:, 0] > 0).long()
    m = keras.Sequential([keras.layers.Dense(2, input_shape=(4,))])
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer="sgd", metrics=["accuracy"])
    m._device = torch.device("cpu")
    m.fit(x, y, batch_size=8, epochs=2, verbose=0, callbacks=[keras.callbacks.TensorBoard(str(tmp_path / "tb"))])
    from tf_yarn_b200.estimator import summary
    sc = summary.read_scalars(str(tmp_path / "tb"))
    assert set(sc["name"]) >= {"epoch_loss", "epoch_accuracy"} and sorted(set(sc["step"])) == [0, 1]


def test_start_tf_board_advertises_its_url_and_reports_failures(monkeypatch, tmp_path):
    from tf_yarn_b200 import tensorboard as tb
    monkeypatch.setenv("TFY_TASK_KEY", "tensorboard:0")
    monkeypatch.setenv("TB_EXTRA_ARGS", "--reload_interval 5")
    launched = {}

    class FakeBoard:
        def configure(self, argv):
            launched["argv"] = argv

        def launch(self):
            launched["launched"] = True

    fake = types.ModuleType("tensorboard.program")
    fake.TensorBoard = FakeBoard
    pkg = types.ModuleType("tensorboard")
    pkg.program = fake
    monkeypatch.setitem(sys.modules, "tensorboard", pkg)
    monkeypatch.setitem(sys.modules, "tensorboard.program", fake)
    client = FakeClient()
    url = tb.start_tf_board(client, str(tmp_path))
    assert url.startswith("http://") and launched["launched"]
    assert launched["argv"][:2] == ["tensorboard", f"--logdir={tmp_path}"] and launched["argv"][-2:] == ["--reload_interval", "5"]
    assert client.kv["tensorboard:0/url"].decode() == url and "tensorboard:0/start" in client.kv.keys()

    def boom(self, argv):
        raise RuntimeError("port in use")
    monkeypatch.setattr(FakeBoard, "configure", boom)
    client2 = FakeClient()
    assert tb.start_tf_board(client2, str(tmp_path)) is None
    assert b"port in use" in client2.kv["tensorboard:0/stop"]
    monkeypatch.setenv("TB_TERMINATION_TIMEOUT_SECONDS", "7")
    assert tb.get_termination_timeout() == 7
