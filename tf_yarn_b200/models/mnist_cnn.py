"""The MNIST CNN of the reference's PyTorch / Horovod-Keras examples.

Conv(1->32, 3) -> Conv(32->64, 3) -> maxpool 2 -> dropout .25 -> FC 9216->128 ->
dropout .5 -> FC 128->10: 1 199 882 parameters (reference:
tf_yarn/examples/pytorch/pytorch_distributed_example.py:44-67; it is the canonical
Horovod ``keras_mnist`` network compiled with ``Adadelta(1.0 * hvd.size())`` and
``sparse_categorical_crossentropy``, reference README.md:104-109).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

N_PARAMS = 1_199_882


def keras_mnist_cnn(logits: bool = True):
    """mini-Keras Sequential version (NHWC input 28x28x1)."""
    from tf_yarn_b200 import keras
    from tf_yarn_b200.keras import layers
    model = keras.Sequential(name="mnist_cnn")
    model.add(layers.Conv2D(32, (3, 3), activation="relu", input_shape=(28, 28, 1)))
    model.add(layers.Conv2D(64, (3, 3), activation="relu"))
    model.add(layers.MaxPooling2D(pool_size=(2, 2)))
    model.add(layers.Dropout(0.25))
    model.add(layers.Flatten())
    model.add(layers.Dense(128, activation="relu"))
    model.add(layers.Dropout(0.5))
    model.add(layers.Dense(10, activation=None if logits else "softmax"))
    return model


class TorchMnistCnn(nn.Module):
    """Plain torch version (NCHW input), as in the reference's PyTorch example."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 32, 3, 1)
        self.conv2 = nn.Conv2d(32, 64, 3, 1)
        self.dropout1 = nn.Dropout(0.25)
        self.dropout2 = nn.Dropout(0.5)
        self.fc1 = nn.Linear(9216, 128)
        self.fc2 = nn.Linear(128, 10)

    def forward(self, x):
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        x = F.max_pool2d(x, 2)
        x = self.dropout1(x)
        x = torch.flatten(x, 1)
        x = F.relu(self.fc1(x))
        x = self.dropout2(x)
        return F.log_softmax(self.fc2(x), dim=1)


def synthetic_mnist(n: int, seed: int = 0, nhwc: bool = True):
    """Random images / labels with the MNIST shape (there is no dataset on an air-gapped box)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((n, 28, 28, 1) if nhwc else (n, 1, 28, 28), generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    return x, y
