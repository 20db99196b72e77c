"""mini-Keras on CPU: building, fit/evaluate/predict, callbacks, persistence, optimizer math."""
import numpy as np
import pytest
import torch

from tf_yarn_b200 import data, keras
from tf_yarn_b200.keras import layers
from tf_yarn_b200.models.mnist_cnn import N_PARAMS, keras_mnist_cnn


def _blobs(n=512, d=11, classes=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    w = torch.randn(d, classes, generator=g)
    return x, (x @ w).argmax(1)


def _mlp(d=11, classes=3):
    m = keras.Sequential()
    m.add(layers.Dense(32, activation="relu", input_shape=(d,)))
    m.add(layers.Dense(classes))
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True),
              optimizer=keras.optimizers.Adam(0.01), metrics=["accuracy"])
    return m


def test_mnist_cnn_has_the_reference_parameter_count():
    m = keras_mnist_cnn()
    assert m.count_params() == N_PARAMS == 1_199_882
    assert m.output_shape == (None, 10)
    lines = []
    m.summary(print_fn=lines.append)
    assert "Total params: 1,199,882" in lines[0]


def test_fit_learns_and_history():
    x, y = _blobs()
    m = _mlp()
    h = m.fit(x, y, batch_size=64, epochs=15, verbose=0)
    assert h.history["loss"][-1] < h.history["loss"][0] * 0.5
    assert h.history["accuracy"][-1] > 0.85
    loss, acc = m.evaluate(x, y, batch_size=128)
    assert acc > 0.85 and loss < h.history["loss"][0]
    assert m.predict(x[:10]).shape == (10, 3)


def test_fit_on_dataset_with_steps_per_epoch():
    x, y = _blobs()
    ds = data.Dataset.from_tensor_slices((x, y)).shuffle(100, seed=1).batch(32).repeat()
    m = _mlp()
    h = m.fit(x=ds, steps_per_epoch=10, epochs=3, verbose=0)
    assert len(h.history["loss"]) == 3


def test_callbacks_checkpoint_lr_schedule_and_batch_hooks(tmp_path):
    x, y = _blobs(256)
    m = _mlp()
    seen = []
    cbs = [keras.callbacks.ModelCheckpoint(str(tmp_path / "ckpt-{epoch}.keras")),
           keras.callbacks.LearningRateScheduler(lambda epoch, lr: lr * 0.5),
           keras.callbacks.LambdaCallback(on_train_batch_end=lambda b, logs: seen.append(logs["loss"]))]
    m.fit(x, y, batch_size=64, epochs=2, verbose=0, callbacks=cbs)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["ckpt-1.keras", "ckpt-2.keras"]
    assert len(seen) == 8 and all(np.isfinite(v) for v in seen)
    assert m.get_learning_rate() == pytest.approx(0.01 * 0.25)


def test_save_load_round_trip(tmp_path):
    x, y = _blobs(256)
    m = _mlp()
    m.fit(x, y, batch_size=64, epochs=3, verbose=0)
    path = str(tmp_path / "model.keras")
    m.save(path)
    m2 = keras.models.load_model(path)
    np.testing.assert_allclose(m.predict(x[:16]), m2.predict(x[:16]), rtol=1e-5, atol=1e-6)
    assert m2.evaluate(x, y, return_dict=True)["accuracy"] == pytest.approx(m.evaluate(x, y, return_dict=True)["accuracy"])
    m.save_weights(str(tmp_path / "w.pt"))
    m3 = _mlp()
    m3.load_weights(str(tmp_path / "w.pt"))
    np.testing.assert_allclose(m.predict(x[:16]), m3.predict(x[:16]), rtol=1e-5, atol=1e-6)


def test_conv_pool_flatten_shapes_follow_keras_nhwc():
    m = keras.Sequential([layers.Conv2D(8, 3, activation="relu", input_shape=(12, 12, 3)),
                          layers.MaxPooling2D(2), layers.Conv2D(16, 3, padding="same"), layers.Flatten(),
                          layers.Dense(4, activation="softmax")])
    m.build()
    assert [ly.output_shape_ for ly in m.layers] == [(10, 10, 8), (5, 5, 8), (5, 5, 16), (400,), (4,)]
    out = m.predict(torch.rand(2, 12, 12, 3))
    assert out.shape == (2, 4) and np.allclose(out.sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("name", ["sgd", "adadelta", "adam", "adagrad"])
def test_optimizer_descriptors_map_to_both_backends(name):
    opt = keras.optimizers.get(name)
    spec = opt.to_spec()
    assert spec.kind.startswith(name[:4])
    p = torch.nn.Parameter(torch.ones(3))
    assert isinstance(opt.to_torch([p]), torch.optim.Optimizer)
    assert keras.optimizers.from_config(opt.get_config()).learning_rate == opt.learning_rate


def test_unknown_identifiers_raise():
    with pytest.raises(ValueError):
        keras.optimizers.get("lion")
    with pytest.raises(ValueError):
        keras.losses.get("hinge-ish")
    with pytest.raises(ValueError):
        layers.Dense(3, activation="swoosh")


def test_fast_path_plan_grammar():
    from tf_yarn_b200.keras import fastpath
    m = keras_mnist_cnn()
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer="adadelta",
              metrics=["accuracy"])
    m.build()
    kinds = [s.kind for s in fastpath.build_plan(m)]
    assert kinds == ["conv", "conv", "flatten", "dense", "head"]
    # softmax output + probability loss is the same fused head
    m2 = keras_mnist_cnn(logits=False)
    m2.compile(loss="sparse_categorical_crossentropy", optimizer="adadelta")
    m2.build()
    assert [s.kind for s in fastpath.build_plan(m2)][-1] == "head"
    # mismatched head (softmax layer feeding a from_logits loss) or another loss: autograd engine
    m3 = keras_mnist_cnn(logits=False)
    m3.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer="adadelta")
    m3.build()
    assert fastpath.build_plan(m3) is None
    m4 = keras_mnist_cnn()
    m4.compile(loss="mse", optimizer="adadelta")
    m4.build()
    assert fastpath.build_plan(m4) is None


def test_keras_estimator_namespace_matches_tf_keras(tmp_path):
    """tf.keras.estimator.model_to_estimator(model, config=RunConfig(model_dir=...)) (reference: examples/keras_example.py:64-65)."""
    from tf_yarn_b200 import estimator as est
    from tf_yarn_b200 import keras as k
    m = k.Sequential([k.layers.Dense(3, input_shape=(4,))])
    m.compile(loss="sparse_categorical_crossentropy", optimizer="sgd", metrics=["accuracy"])
    e = k.estimator.model_to_estimator(m, config=est.RunConfig(model_dir=str(tmp_path)))
    assert isinstance(e, est.Estimator) and e.model_dir == str(tmp_path)
