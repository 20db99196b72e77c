"""Model-level API of the mini-Keras facade on CPU: the calls the reference's examples and docs make
(examples/keras_example.py:55-62, native_keras_with_gloo_example.py:65-90, README.md:104-113)."""
import numpy as np
import pytest
import torch

from tf_yarn_b200 import data, keras


def _xy(n=96, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 4, generator=g)
    return x, (x[:, 0] + x[:, 1] > 0).long()


def _model():
    m = keras.Sequential()
    m.add(keras.layers.Dense(8, activation="relu", input_shape=(4,)))
    m.add(keras.layers.Dense(2))
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer=keras.optimizers.SGD(0.2),
              metrics=["accuracy"])
    m._device = torch.device("cpu")
    return m


def test_fit_with_validation_data_callables_and_history():
    x, y = _xy()
    xv, yv = _xy(32, seed=1)
    m = _model()
    h = m.fit(lambda: x, lambda: y, batch_size=16, epochs=6, verbose=1, validation_data=(xv, yv))
    assert set(h.history) == {"loss", "accuracy", "val_loss", "val_accuracy"} and len(h.history["loss"]) == 6
    assert h.history["loss"][-1] < h.history["loss"][0] and h.history["val_accuracy"][-1] > 0.7
    loss, acc = m.evaluate(xv, yv, batch_size=8)
    assert acc == pytest.approx(h.history["val_accuracy"][-1], abs=1e-6) and loss > 0
    assert m.evaluate(xv, yv, return_dict=True).keys() == {"loss", "accuracy"}


def test_fit_on_a_dataset_with_steps_per_epoch_and_stop_training():
    x, y = _xy()
    ds = data.Dataset.from_tensor_slices((x, y)).shuffle(96, seed=0).batch(16).repeat()
    m = _model()
    seen = []
    cb = keras.callbacks.LambdaCallback(on_epoch_end=lambda e, logs: seen.append(e))
    m.fit(ds, epochs=3, steps_per_epoch=4, verbose=0, callbacks=[cb])
    assert seen == [0, 1, 2]

    class StopAfterFirst(keras.callbacks.Callback):
        def on_epoch_end(self, epoch, logs=None):
            self.model.stop_training = True

    h = _model().fit(x, y, batch_size=16, epochs=5, verbose=0, callbacks=[StopAfterFirst()])
    assert len(h.history["loss"]) == 1
    finite = data.Dataset.from_tensor_slices((x, y)).batch(32)          # cardinality known: no steps_per_epoch needed
    h = _model().fit(finite, epochs=2, verbose=0)
    assert len(h.history["loss"]) == 2


def test_predict_call_weights_and_summary(tmp_path):
    x, y = _xy(40)
    m = _model()
    m.fit(x, y, batch_size=8, epochs=1, verbose=0)
    p = m.predict(x, batch_size=16)
    assert isinstance(p, np.ndarray) and p.shape == (40, 2)
    assert np.allclose(p, m(x).detach().numpy(), atol=1e-6)
    assert np.allclose(m.predict(x, batch_size=16, steps=1), p[:16])
    w = m.get_weights()
    assert [a.shape for a in w] == [(8, 4), (8,), (2, 8), (2,)]
    m2 = _model()
    m2.set_weights(w)
    assert np.allclose(m2.predict(x), p, atol=1e-6)
    with pytest.raises(ValueError, match="expects 4 arrays"):
        m2.set_weights(w[:2])
    with pytest.raises(ValueError, match="shape mismatch"):
        m2.set_weights([w[0].T] + w[1:])
    m.save_weights(str(tmp_path / "w.pt"))
    m3 = _model()
    m3.load_weights(str(tmp_path / "w.pt"))
    assert np.allclose(m3.predict(x), p, atol=1e-6)
    lines = []
    m.summary(print_fn=lines.append)
    assert "dense (Dense)" in lines[0] and "dense_1 (Dense)" in lines[0] and "Total params: 58" in lines[0]
    assert m.output_shape == (None, 2) and m.count_params() == 58
    with pytest.raises(RuntimeError, match="after the model was built"):
        m.add(keras.layers.Dense(1))


def test_from_torch_wraps_an_arbitrary_module_and_errors_are_explicit():
    net = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.Tanh(), torch.nn.Linear(6, 2))
    m = keras.Model.from_torch(net, input_shape=(4,))
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer="adam")
    m._device = torch.device("cpu")
    x, y = _xy()
    h = m.fit(x, y, batch_size=16, epochs=4, verbose=0)
    assert h.history["loss"][-1] < h.history["loss"][0]
    with pytest.raises(ValueError, match="no layers"):
        keras.Sequential().build()
    with pytest.raises(ValueError, match="input_shape"):
        keras.Sequential([keras.layers.Dense(2)]).build()
    uncompiled = keras.Sequential([keras.layers.Dense(2, input_shape=(4,))])
    with pytest.raises(RuntimeError, match="compile"):
        uncompiled.fit(x, y, verbose=0)
    with pytest.raises(ValueError, match="unknown optimizer"):
        uncompiled.compile(optimizer="rmsprop-like", loss="mse")
    with pytest.raises(TypeError, match="unexpected arguments"):
        keras.layers.Dense(2, kernel_regularizer="l2")


def test_load_model_resumes_with_the_saved_optimizer_state(tmp_path, caplog):
    """save() -> load_model() -> fit() continues exactly like an uninterrupted run (Adam moments and step count are
    part of the checkpoint), which is what the reference's evaluator / resume flows rely on with tf.keras."""
    x, y = _xy(64)

    def fresh():
        torch.manual_seed(5)
        m = keras.Sequential([keras.layers.Dense(6, activation="tanh", input_shape=(4,)), keras.layers.Dense(2)])
        m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer=keras.optimizers.Adam(0.05))
        m._device = torch.device("cpu")
        return m

    straight = fresh()
    straight.fit(x, y, batch_size=16, epochs=3, shuffle=False, verbose=0)
    first = fresh()
    first.fit(x, y, batch_size=16, epochs=2, shuffle=False, verbose=0)
    path = str(tmp_path / "ckpt-2")
    first.save(path)
    resumed = keras.models.load_model(path)
    resumed._device = torch.device("cpu")
    assert resumed._pending_engine_state["kind"] == "eager"
    resumed.fit(x, y, batch_size=16, epochs=1, shuffle=False, verbose=0)
    for a, b in zip(resumed.get_weights(), straight.get_weights()):
        assert np.allclose(a, b, atol=1e-6)
    assert resumed._pending_engine_state is None

    weights_only = keras.models.load_model(path)                 # without the optimizer state the run diverges
    weights_only._device = torch.device("cpu")
    weights_only._pending_engine_state = None
    weights_only.fit(x, y, batch_size=16, epochs=1, shuffle=False, verbose=0)
    assert not all(np.allclose(a, b, atol=1e-6) for a, b in zip(weights_only.get_weights(), straight.get_weights()))

    foreign = keras.models.load_model(path)
    foreign._device = torch.device("cpu")
    foreign._pending_engine_state = {"kind": "fused", "master": torch.zeros(3)}
    with caplog.at_level("WARNING"):
        foreign.fit(x, y, batch_size=16, epochs=1, shuffle=False, verbose=0)
    assert "optimizer starts fresh" in caplog.text


def test_single_rank_readers_never_start_a_collective_when_state_is_sharded(tmp_path):
    """On several GPUs the fp32 master / optimizer state is sharded; ModelCheckpoint runs on the chief ALONE, so
    save() / get_weights() must not call the engine's gathering (collective) paths: they read the replicated
    parameters and leave the optimizer state out."""
    from types import SimpleNamespace
    from tf_yarn_b200.keras.engine import GraphTrainEngine
    m = _model()
    m.build()

    class Boom:
        def __getattr__(self, name):
            raise AssertionError(f"collective path touched: {name}")

    eng = GraphTrainEngine.__new__(GraphTrainEngine)
    eng.comm = SimpleNamespace(world=8)
    eng.fused = Boom()                       # gather_state / master_tensors would go through here
    m._engine = eng
    assert m._sharded_across_ranks()
    w = m.get_weights()
    assert [a.shape for a in w] == [(8, 4), (8,), (2, 8), (2,)]
    path = str(tmp_path / "ck")
    m.save(path)
    import pickle
    payload = pickle.load(open(path, "rb"))
    assert payload["optimizer_state"] is None and set(payload["weights"]) == set(m.net.state_dict())
    eng.comm = SimpleNamespace(world=1)
    assert not m._sharded_across_ranks()


def test_evaluate_accepts_a_pair_of_arrays_and_dict_outputs():
    """`validation_data_fn=lambda: (x_val, y_val)` (the README quick start) reaches the evaluator as ONE argument; and
    models whose outputs are dicts (BERT's heads) are evaluated with a loss over dicts."""
    x, y = _xy(48)
    m = _model()
    m.fit(x, y, batch_size=16, epochs=1, verbose=0)
    direct = m.evaluate(x, y, return_dict=True)
    assert m.evaluate((x, y), return_dict=True) == direct
    assert m.evaluate(lambda: (x, y), return_dict=True) == direct
    batches = [(x[:24], y[:24]), (x[24:], y[24:])]                        # an iterable of two BATCHES stays an iterable
    assert m.evaluate(batches, return_dict=True)["accuracy"] == pytest.approx(direct["accuracy"])

    class TwoHeads(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = torch.nn.Linear(4, 2), torch.nn.Linear(4, 3)

        def forward(self, feats):
            return {"a": self.a(feats["x"]), "b": self.b(feats["x"])}

    def loss(targets, out):
        return torch.nn.functional.cross_entropy(out["a"], targets["ya"]) + \
            torch.nn.functional.cross_entropy(out["b"], targets["yb"])
    d = keras.Model.from_torch(TwoHeads())
    d.compile(loss=loss, optimizer="sgd")
    d._device = torch.device("cpu")
    data_ = [({"x": x[:16]}, {"ya": y[:16], "yb": y[:16]}), ({"x": x[16:32]}, {"ya": y[16:32], "yb": y[16:32]})]
    d.fit(iter(data_ * 3), steps_per_epoch=6, epochs=1, verbose=0)
    res = d.evaluate(iter(data_), return_dict=True)
    assert set(res) == {"loss"} and res["loss"] > 0


def test_validation_split_initializers_and_ignored_fit_arguments(caplog):
    x, y = _xy(80)
    m = _model()
    with caplog.at_level("WARNING"):
        h = m.fit(x, y, batch_size=16, epochs=2, verbose=0, validation_split=0.25, class_weight={0: 1.0})
    assert "class_weight" in caplog.text                                   # not silently dropped
    assert set(h.history) == {"loss", "accuracy", "val_loss", "val_accuracy"}
    held_out = m.evaluate(x[60:], y[60:], return_dict=True)                # the LAST 25 % were the validation set
    assert held_out["loss"] == pytest.approx(h.history["val_loss"][-1], rel=1e-5)
    with pytest.raises(ValueError, match="validation_split"):
        _model().fit(iter([(x, y)]), epochs=1, verbose=0, validation_split=0.2, steps_per_epoch=1)

    torch.manual_seed(0)
    d = keras.layers.Dense(64, input_shape=(128,), kernel_initializer="he_normal", bias_initializer="ones")
    d.build((128,))
    assert float(d.module.bias.detach().min()) == 1.0
    assert abs(float(d.module.weight.detach().std()) - (2 / 128) ** 0.5) < 0.02
    z = keras.layers.Dense(4, kernel_initializer="zeros")
    z.build((3,))
    assert float(z.module.weight.detach().abs().max()) == 0.0
    c = keras.layers.Conv2D(8, 3, kernel_initializer=lambda w: torch.nn.init.constant_(w, 0.5))
    c.build((6, 6, 2))
    assert float(c.module.weight.detach().min()) == 0.5
    with pytest.raises(ValueError, match="unknown initializer"):
        keras.layers.Dense(2, kernel_initializer="orthogonal-ish").build((3,))
