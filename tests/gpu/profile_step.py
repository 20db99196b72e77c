"""Run a few captured MNIST-CNN train steps between cudaProfilerStart/Stop (for ncu launch lists).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tests/gpu/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_yarn_b200 import hvd, keras  # noqa: E402
from tf_yarn_b200.models.mnist_cnn import keras_mnist_cnn, synthetic_mnist  # noqa: E402


def main():
    steps = int(os.environ.get("PROFILE_STEPS", "3"))
    torch.cuda.set_device(0)
    hvd.init()
    model = keras_mnist_cnn()
    model.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True),
                  optimizer=hvd.DistributedOptimizer(keras.optimizers.Adadelta(1.0)))
    x, y = synthetic_mnist(128 * 8)
    model.fit(x, y, batch_size=128, epochs=1, shuffle=False, verbose=0)     # builds + captures
    eng = model._engine
    xd, yd = x.cuda(), y.cuda()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(steps):
        t = eng.stage_inputs(xd[i * 128:(i + 1) * 128], yd[i * 128:(i + 1) * 128])
        eng.launch_step(t)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled", steps, "steps")


if __name__ == "__main__":
    main()
