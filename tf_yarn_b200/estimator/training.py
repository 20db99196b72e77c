"""``train_and_evaluate``: role-dependent behaviour driven by ``TF_CONFIG``.

* chief / worker : ``estimator.train`` (parameter-server data plane when the cluster has ps tasks);
* ps             : serve the shard forever (the task leaves through the stop barrier);
* evaluator      : evaluate every new checkpoint until ``max_steps`` was evaluated;
* no TF_CONFIG   : train, then evaluate once (local mode).

(reference call site: tf_yarn/tensorflow/tasks/tf_task_common.py:46-50; TF_CONFIG is exported by
tf_yarn/tensorflow/cluster.py:41-50.)
"""
from __future__ import annotations

import logging
import os
import time
from datetime import datetime, timedelta
from typing import Optional

from tf_yarn_b200.estimator import checkpoint as ckpt
from tf_yarn_b200.estimator.config import ClusterInfo
from tf_yarn_b200.estimator.spec import EvalSpec, GraphKeys, TrainSpec

logger = logging.getLogger(__name__)


def continuous_eval(estimator, train_spec: TrainSpec, eval_spec: EvalSpec, timeout_secs: Optional[float] = None,
                    stop_cond=None, evaluated_steps: Optional[set] = None) -> Optional[dict]:
    """Evaluate checkpoints as they appear; stops once ``max_steps`` was evaluated, on idle timeout
    (counted from the last evaluated checkpoint) or when ``stop_cond()`` is true."""
    evaluated = set(evaluated_steps or ())
    last_result = None
    if eval_spec.start_delay_secs:
        time.sleep(eval_spec.start_delay_secs)
    stamp = datetime.now()
    done = False
    exporters = eval_spec.exporters or []
    if not isinstance(exporters, (list, tuple)):
        exporters = [exporters]
    while not done:
        if stop_cond is not None and stop_cond():
            break
        if timeout_secs and datetime.now() > stamp + timedelta(seconds=timeout_secs):
            logger.info("Stopping evaluation due to timeout")
            break
        st = ckpt.get_checkpoint_state(estimator.model_dir)
        todo = [p for p in (st.all_model_checkpoint_paths if st else [])
                if ckpt.step_of(p) not in evaluated and os.path.exists(p)]
        for path in sorted(todo, key=ckpt.step_of):
            stamp = datetime.now()
            try:
                result = estimator.evaluate(eval_spec.input_fn, steps=eval_spec.steps, hooks=eval_spec.hooks,
                                            name=eval_spec.name, checkpoint_path=path)
            except FileNotFoundError:               # pruned by the chief between listing and loading: skip it
                logger.info("checkpoint %s disappeared before it could be evaluated", path)
                evaluated.add(ckpt.step_of(path))
                continue
            last_result = result
            evaluated.add(ckpt.step_of(path))
            gs = last_result.get(GraphKeys.GLOBAL_STEP) if last_result else None
            if train_spec.max_steps and gs is not None and gs >= train_spec.max_steps:
                done = True
            for exporter in exporters:
                exporter.export(estimator, os.path.join(estimator.model_dir, "export", exporter.name), path,
                                last_result, done)
            if done:
                break
        if not done:
            time.sleep(min(eval_spec.throttle_secs, 1.0) if not todo else 0.0)
            if not todo and eval_spec.throttle_secs > 1.0:
                time.sleep(eval_spec.throttle_secs - 1.0)
    return last_result


def train_and_evaluate(estimator, train_spec: TrainSpec, eval_spec: EvalSpec):
    cluster = ClusterInfo.from_env()
    role = cluster.task_type
    if cluster.distributed and role == "ps":
        import torch
        if torch.cuda.is_available() and os.environ.get("TFY_GPU_IDS") and \
                os.environ.get("TFY_PS_PLANE", "auto") != "shm":
            from tf_yarn_b200.estimator import ps_hbm
            ps_hbm.serve(cluster)  # never returns
        from tf_yarn_b200.estimator import ps
        ps.serve(cluster)          # never returns
        return None
    if cluster.distributed and role == "evaluator":
        return continuous_eval(estimator, train_spec, eval_spec, timeout_secs=1200), None
    estimator.train(train_spec.input_fn, hooks=list(train_spec.hooks), max_steps=train_spec.max_steps)
    if not cluster.distributed:
        result = estimator.evaluate(eval_spec.input_fn, steps=eval_spec.steps, hooks=list(eval_spec.hooks),
                                    name=eval_spec.name)
        exporters = eval_spec.exporters or []
        if not isinstance(exporters, (list, tuple)):
            exporters = [exporters]
        for exporter in exporters:
            exporter.export(estimator, os.path.join(estimator.model_dir, "export", exporter.name),
                            estimator.latest_checkpoint(), result, True)
        return result, None
    return None
