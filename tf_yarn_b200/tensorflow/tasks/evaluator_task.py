"""Evaluator side task: evaluate every new checkpoint of the training job.

(reference: tf_yarn/tensorflow/tasks/evaluator_task.py:18-158)

Estimator experiments: resume the set of already evaluated steps from the eval event
file, sleep ``start_delay_secs``, evaluate each not-yet-seen ``model.ckpt-N`` (+ exporters),
stop when ``max_steps`` was evaluated or after the idle timeout.

Keras experiments: ``load_model(ckpt)`` + ``model.evaluate(validation_data_fn())`` for each
new checkpoint file in ``model_dir``.  The reference's Keras loop only ends by timeout; here
it also ends once every training task has published ``stop`` and nothing is left to evaluate.
"""
from __future__ import annotations

import logging
import os
import re
import time
from datetime import datetime, timedelta
from typing import Callable, Set

from tf_yarn_b200 import _task_commons, event
from tf_yarn_b200._task_commons import TaskClient, get_task, get_task_key, setup_logging
from tf_yarn_b200.estimator import checkpoint as ckpt
from tf_yarn_b200.estimator import continuous_eval
from tf_yarn_b200.estimator import summary as summary_lib
from tf_yarn_b200.tensorflow import Experiment, KerasExperiment, metrics

logger = logging.getLogger(__name__)

KERAS_POLL_SECS = float(os.environ.get("TFY_KERAS_EVAL_POLL_SECS", "30"))
EVAL_TIMEOUT_SECS = float(os.environ.get("TFY_EVAL_TIMEOUT_SECS", "1200"))     # 20 min idle timeout


def evaluator_fn(client) -> None:
    experiment = _task_commons._get_experiment(client)
    if isinstance(experiment, Experiment):
        evaluate(experiment, timeout_in_secs=EVAL_TIMEOUT_SECS)
    elif isinstance(experiment, KerasExperiment):
        keras_evaluate(experiment, stop_cond=_training_finished(client), timeout_in_secs=EVAL_TIMEOUT_SECS,
                       client=client)
    else:
        raise ValueError("experiment must be an Experiment or a KerasExperiment")


def _training_finished(client) -> Callable[[], bool]:
    """True once every task of the training cluster has published its ``stop`` event."""
    tasks = _task_commons._get_cluster_tasks(client)

    def cond() -> bool:
        return all(client.kv.get(f"{t.to_container_key().to_kv_str()}/stop") is not None for t in tasks)
    return cond


def stop_cond_reached(stop_cond, timeout_in_secs, timestamp) -> bool:
    if stop_cond and stop_cond():
        logger.info("Stop condition met")
        return True
    if timeout_in_secs and datetime.now() > (timestamp + timedelta(seconds=timeout_in_secs)):
        logger.info("Stopping evaluation due to timeout")
        return True
    return False


def _get_step(checkpoint: str) -> int:
    if "model.ckpt-" in checkpoint:
        return int(checkpoint.split("model.ckpt-")[1])
    # Keras ModelCheckpoint names ("weights.02.h5", "model-7.pt", "ckpt_0003.keras"): the LAST integer of the
    # stem, extension(s) stripped -- never a digit of the extension itself
    stem = os.path.basename(checkpoint)
    while True:
        root, ext = os.path.splitext(stem)
        if not ext or not root or re.fullmatch(r"\.\d+", ext):
            break
        stem = root
    # "weights.02-0.35.h5" ({epoch:02d}-{val_loss:.2f}): a metric formatted as a float is not the step
    m = re.search(r"(\d+)(?!.*\d)", re.sub(r"\d+\.\d+", "", stem)) or re.search(r"(\d+)(?!.*\d)", stem)
    if m is None:
        raise ValueError(f"cannot parse a training step out of checkpoint name {checkpoint!r}: use a "
                         "ModelCheckpoint filepath with {epoch} or a step number in it")
    return int(m.group(1))


def _get_all_checkpoints(model_dir: str):
    st = ckpt.get_checkpoint_state(model_dir)
    return st.all_model_checkpoint_paths if st else []


def _keras_checkpoints(model_dir: str):
    """Checkpoint files written by ModelCheckpoint (anything that is not an event file / temp file)."""
    if not os.path.isdir(model_dir):
        return []
    out = []
    for name in sorted(os.listdir(model_dir)):
        path = os.path.join(model_dir, name)
        if os.path.isfile(path) and not summary_lib.is_event_file(name) and ".tmp" not in name \
                and name != ckpt.INDEX_FILE:
            out.append(path)
    return out


def get_ckpt_to_eval(model_dir: str, evaluated_checkpoints: Set[int]):
    return {c for c in _get_all_checkpoints(model_dir) if _get_step(c) not in evaluated_checkpoints}


def _get_evaluated_checkpoint(eval_dir: str) -> Set[int]:
    return set(metrics.get_all_metrics(eval_dir)["step"])


def get_initial_evaluated_checkpoints(eval_dir: str) -> Set[int]:
    if not os.path.exists(eval_dir) or not os.listdir(eval_dir):
        return set()
    return _get_evaluated_checkpoint(eval_dir)


def keras_evaluate(experiment: KerasExperiment, stop_cond=None, timeout_in_secs=None, client=None) -> None:
    from tf_yarn_b200 import keras
    eval_dir = os.path.join(experiment.model_dir, "eval")
    evaluated = get_initial_evaluated_checkpoints(eval_dir)
    evaluated_before = set(evaluated)
    seen_files: Set = set()
    timestamp = datetime.now()
    writer = None
    n_done = 0
    while True:
        finished = stop_cond_reached(stop_cond, timeout_in_secs, timestamp)
        # dedupe on (path, mtime): two files whose names parse to the same number are both evaluated, and a
        # checkpoint rewritten in place is evaluated again; the parsed number is only the TensorBoard step.
        # Steps found in an existing eval directory (a restarted evaluator) are still skipped.
        todo = []
        for c in _keras_checkpoints(experiment.model_dir):
            key = (c, os.path.getmtime(c))
            if key in seen_files or _get_step(c) in evaluated_before:
                continue
            todo.append(c)
            seen_files.add(key)
        for path in todo:
            timestamp = datetime.now()
            logger.info("Evaluating checkpoint %s", path)
            model = keras.models.load_model(path)
            t0 = time.time()
            result = model.evaluate(experiment.validation_data_fn(), return_dict=True)
            step = _get_step(path)
            evaluated.add(step)
            n_done += 1
            if writer is None:
                writer = summary_lib.writer(eval_dir)
            for k, v in result.items():
                writer.add_scalar(k, v, step)
            writer.flush()
            logger.info("checkpoint %s: %s", os.path.basename(path), result)
            if client is not None:
                task = get_task()
                event.broadcast(client, f"{task}/nb_eval_steps", str(n_done))
                event.broadcast(client, f"{task}/last_training_step", str(step))
                event.broadcast(client, f"{task}/eval_step_mean_duration", str(time.time() - t0))
        if finished and not todo:
            break
        if not todo:
            logger.info("No checkpoint to evaluate; sleeping %.0f s", KERAS_POLL_SECS)
            time.sleep(KERAS_POLL_SECS)
    if writer is not None:
        writer.close()


def evaluate(experiment: Experiment, stop_cond=None, timeout_in_secs=None):
    eval_dir = os.path.join(experiment.estimator.model_dir, "eval")
    evaluated = get_initial_evaluated_checkpoints(eval_dir)
    if evaluated:
        last = max(evaluated)
        logger.info("Last evaluated checkpoint: %s", last)
        if experiment.train_spec.max_steps and last == experiment.train_spec.max_steps:
            logger.info("All checkpoints already evaluated; Max steps: %s", experiment.train_spec.max_steps)
            return None
    return continuous_eval(experiment.estimator, experiment.train_spec, experiment.eval_spec,
                           timeout_secs=timeout_in_secs, stop_cond=stop_cond, evaluated_steps=evaluated)


def main() -> None:
    setup_logging()
    client = TaskClient.from_current()
    task_key = get_task_key()
    task = task_key.to_kv_str()
    event.init_event(client, task, "127.0.0.1:0")
    _task_commons._setup_container_logs(client)
    error = None
    try:
        if task_key.type == "evaluator":
            event.start_event(client, task)
            event.broadcast_train_eval_start_timer(client, task)
            evaluator_fn(client)
            event.broadcast_train_eval_stop_timer(client, task)
        else:
            logger.info("%s: nothing to do", task)
    except Exception as exc:  # noqa: BLE001
        error = exc
    event.stop_event(client, task, error)
    event.broadcast_container_stop_time(client, task)
    if error is not None:
        raise error


if __name__ == "__main__":
    main()
