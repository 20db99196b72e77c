"""Every mini-Keras layer against plain torch with Keras' channels-last semantics, plus config round trips."""
import torch
import torch.nn.functional as F

from tf_yarn_b200 import keras
from tf_yarn_b200.keras import layers as L


def _image_model(*mid, shape=(9, 9, 3)):
    m = keras.Sequential([L.InputLayer(shape)] + list(mid))
    m._device = torch.device("cpu")
    m.build(shape)
    return m


def _nhwc_apply(fn, x):                       # run an NCHW torch op on an NHWC tensor
    return fn(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)


def test_pooling_layers_match_tf_same_and_valid_semantics():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 9, 9, 3, generator=g)
    valid = _image_model(L.MaxPooling2D((2, 2)))
    assert torch.allclose(valid.net(x), _nhwc_apply(lambda t: F.max_pool2d(t, 2, 2), x))
    assert valid.layers[-1].output_shape_ == (4, 4, 3)

    same = _image_model(L.MaxPooling2D((2, 2), padding="same"))
    out = same.net(x)
    assert out.shape == (2, 5, 5, 3) == (2,) + same.layers[-1].output_shape_
    ref = _nhwc_apply(lambda t: F.max_pool2d(F.pad(t, (0, 1, 0, 1), value=float("-inf")), 2, 2), x)   # TF pads at the end
    assert torch.allclose(out, ref)

    avg_valid = _image_model(L.AveragePooling2D((3, 3), strides=(2, 2)))
    assert torch.allclose(avg_valid.net(x), _nhwc_apply(lambda t: F.avg_pool2d(t, 3, 2), x), atol=1e-6)

    avg_same = _image_model(L.AveragePooling2D((2, 2), padding="same"))
    out = avg_same.net(x)
    assert out.shape == (2, 5, 5, 3)
    # the last row / column windows contain ONE real row / column: TF averages over the real elements only
    assert torch.allclose(out[:, :4, :4], _nhwc_apply(lambda t: F.avg_pool2d(t, 2, 2), x), atol=1e-6)
    assert torch.allclose(out[:, 4, 4], x[:, 8, 8], atol=1e-6)
    assert torch.allclose(out[:, 4, 0], x[:, 8, 0:2].mean(dim=1), atol=1e-6)

    gap = _image_model(L.GlobalAveragePooling2D())
    assert torch.allclose(gap.net(x), x.mean(dim=(1, 2)), atol=1e-6) and gap.layers[-1].output_shape_ == (3,)


def test_conv_flatten_norm_activation_embedding_layers():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 8, 8, 3, generator=g)
    m = _image_model(L.Conv2D(4, 3, padding="same", activation="relu"), L.BatchNormalization(),
                     L.LayerNormalization(), L.Activation("tanh"), L.ReLU(), L.Flatten(), L.Dense(5), L.Softmax(),
                     shape=(8, 8, 3))
    m.net.eval()
    out = m.net(x)
    assert out.shape == (2, 5) and torch.allclose(out.sum(-1), torch.ones(2), atol=1e-5)
    conv, bn, ln = m.layers[1].module, m.layers[2].module, m.layers[3].module
    with torch.no_grad():
        h = F.relu(F.conv2d(x.permute(0, 3, 1, 2), conv.weight, conv.bias, padding=1))
        h = bn(h)                                                         # eval mode: running statistics
        h = F.layer_norm(h.permute(0, 2, 3, 1), (4,), ln.weight, ln.bias, ln.eps)      # channel axis, NHWC
        h = F.relu(torch.tanh(h)).reshape(2, -1)                          # Keras flattens in (H, W, C) order
        dense = m.layers[7].module
        ref = F.softmax(F.linear(h, dense.weight, dense.bias), dim=-1)
    assert torch.allclose(out, ref, atol=1e-5)
    assert m.layers[1].output_shape_ == (8, 8, 4) and m.layers[6].output_shape_ == (256,)

    strided = _image_model(L.Conv2D(2, (3, 3), strides=(2, 2)), shape=(9, 9, 3))
    assert strided.layers[-1].output_shape_ == (4, 4, 2) and strided.net(torch.randn(1, 9, 9, 3)).shape == (1, 4, 4, 2)

    emb = keras.Sequential([L.Embedding(10, 4, input_shape=(3,)), L.Flatten(), L.Dense(2)])
    emb._device = torch.device("cpu")
    emb.build((3,))
    assert emb.net(torch.tensor([[1, 2, 3], [4, 5, 6]])).shape == (2, 2)

    drop = L.Dropout(0.5)
    t = torch.ones(1000)
    assert torch.equal(drop.call(t, training=False), t) and 300 < int((drop.call(t, training=True) == 0).sum()) < 700


def test_layer_configs_survive_save_and_load(tmp_path):
    m = keras.Sequential([L.Conv2D(4, 3, activation="relu", input_shape=(8, 8, 1)), L.MaxPooling2D(2, padding="same"),
                          L.AveragePooling2D(2), L.BatchNormalization(momentum=0.9), L.Dropout(0.25), L.Flatten(),
                          L.Dense(6, activation="tanh", use_bias=False), L.LayerNormalization(epsilon=1e-4),
                          L.Activation("gelu"), L.Dense(3)])
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer=keras.optimizers.Adam(1e-3),
              metrics=["accuracy"])
    m._device = torch.device("cpu")
    m.build()
    path = str(tmp_path / "model.ckpt")
    m.save(path)
    m2 = keras.models.load_model(path)
    assert [type(ly).__name__ for ly in m2.layers] == [type(ly).__name__ for ly in m.layers]
    assert m2.layers[1].padding == "same" and m2.layers[3].momentum == 0.9 and m2.layers[6].use_bias is False
    assert m2.layers[7].epsilon == 1e-4 and m2.layers[8].activation_name == "gelu" and m2.layers[4].rate == 0.25
    x = torch.randn(3, 8, 8, 1)
    m.net.eval(), m2.net.eval()
    assert torch.allclose(m.net(x), m2.net.to("cpu")(x), atol=1e-6)
    assert m.count_params() == m2.count_params() > 0
