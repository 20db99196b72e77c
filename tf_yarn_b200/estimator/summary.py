"""TensorBoard event files: writer + iterator (``tf.compat.v1.train.summary_iterator`` stand-in)."""
from __future__ import annotations

import os
from typing import Dict, Iterator, List


def writer(logdir: str):
    from torch.utils.tensorboard import SummaryWriter
    os.makedirs(logdir, exist_ok=True)
    return SummaryWriter(logdir)


def is_event_file(filename: str) -> bool:
    return os.path.basename(filename).startswith("events.out")


def summary_iterator(path: str) -> Iterator:
    """Yield the ``Event`` protos of one event file (fields: ``step``, ``summary.value[].tag/simple_value``)."""
    from tensorboard.backend.event_processing.event_file_loader import LegacyEventFileLoader
    yield from LegacyEventFileLoader(path).Load()


def read_scalars(model_path: str) -> Dict[str, List]:
    """All scalar summaries under ``model_path`` as ``{'step': [...], 'name': [...], 'value': [...]}``."""
    frame: Dict[str, List] = {"step": [], "name": [], "value": []}
    if not os.path.isdir(model_path):
        return frame
    for name in sorted(os.listdir(model_path)):
        if not is_event_file(name):
            continue
        for ev in summary_iterator(os.path.join(model_path, name)):
            if not ev.HasField("summary"):
                continue
            for value in ev.summary.value:
                val = None
                if value.HasField("simple_value"):
                    val = value.simple_value
                elif value.HasField("tensor") and value.tensor.float_val:
                    val = value.tensor.float_val[0]
                elif value.HasField("tensor") and value.tensor.tensor_content:
                    import numpy as np
                    arr = np.frombuffer(value.tensor.tensor_content, dtype=np.float32)
                    val = float(arr[0]) if arr.size else None
                if val is not None:
                    frame["step"].append(ev.step)
                    frame["name"].append(value.tag)
                    frame["value"].append(val)
    return frame
