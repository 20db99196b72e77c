"""Step watchdog (utils/watchdog.py): a training loop that stops making progress fails the task instead of hanging."""
import subprocess
import sys
import time

import torch

from tf_yarn_b200 import keras
from tf_yarn_b200.utils import watchdog


def test_fires_without_heartbeats_and_not_with_them():
    fired = []
    wd = watchdog.StepWatchdog(0.6, "loop", action=lambda what, idle: fired.append((what, idle))).start()
    t0 = time.time()
    while time.time() - t0 < 1.0:                 # beating: never fires
        wd.beat()
        time.sleep(0.02)
    assert not fired and not wd.fired
    time.sleep(1.5)                               # silent: fires once
    wd.close()
    assert len(fired) == 1 and fired[0][0] == "loop" and fired[0][1] > 0.6 and wd.fired


def test_zero_timeout_is_inert_and_env_overrides_the_default(monkeypatch):
    wd = watchdog.StepWatchdog(0.0).start()
    assert wd._thread is None
    wd.close()
    monkeypatch.delenv("TFY_STEP_TIMEOUT_SECS", raising=False)
    assert watchdog.default_timeout(False) == 0.0
    assert watchdog.default_timeout(True) == watchdog.DEFAULT_DISTRIBUTED_SECS
    monkeypatch.setenv("TFY_STEP_TIMEOUT_SECS", "12.5")
    assert watchdog.default_timeout(False) == 12.5 and watchdog.default_timeout(True) == 12.5
    monkeypatch.setenv("TFY_STEP_TIMEOUT_SECS", "0")
    assert watchdog.default_timeout(True) == 0.0


def test_fit_beats_every_step_and_a_stalled_step_trips_the_watchdog(monkeypatch):
    fired = []
    monkeypatch.setattr(watchdog, "_default_action", lambda what, idle: fired.append(what))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 4, generator=g)
    y = (x[:, 0] > 0).long()

    def model():
        m = keras.Sequential([keras.layers.Dense(2, input_shape=(4,))])
        m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer="sgd")
        m._device = torch.device("cpu")
        return m

    monkeypatch.setenv("TFY_STEP_TIMEOUT_SECS", "0")
    model().fit(x, y, batch_size=8, epochs=1, verbose=0)          # one-time warm-up of the CPU engine, unwatched
    monkeypatch.setenv("TFY_STEP_TIMEOUT_SECS", "2.0")
    model().fit(x, y, batch_size=8, epochs=3, verbose=0)          # 24 quick steps: no alarm
    assert fired == []
    stall = keras.callbacks.LambdaCallback(on_train_batch_end=lambda b, logs: time.sleep(3.5) if b == 2 else None)
    m = model()
    m.fit(x, y, batch_size=8, epochs=1, verbose=0, callbacks=[stall])
    assert fired == ["Model.fit"]
    assert m._watchdog is None                                     # released when fit returns


def test_default_action_ends_the_process_with_the_retry_code():
    code = ("import time\n"
            "from tf_yarn_b200.utils import watchdog\n"
            "wd = watchdog.StepWatchdog(0.2, 'stuck loop').start()\n"
            "time.sleep(30)\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert res.returncode == watchdog.EXIT_CODE
    assert "stuck loop made no progress" in res.stderr and "Thread" in res.stderr     # message + stack dump
