"""Keras experiment descriptor (reference: tf_yarn/tensorflow/keras_experiment.py:5-11)."""
from typing import Any, Callable, Dict, NamedTuple, Optional


class KerasExperiment(NamedTuple):
    model: Any                                   # compiled tf_yarn_b200.keras.Model
    model_dir: str                               # where ModelCheckpoint writes / the evaluator reads
    train_params: Dict[str, Any]                 # kwargs of model.fit (steps_per_epoch, callbacks, epochs, ...)
    input_data_fn: Optional[Callable] = None     # -> x of model.fit (tensor or Dataset)
    target_data_fn: Optional[Callable] = None    # -> y of model.fit
    validation_data_fn: Optional[Callable] = None  # -> data the evaluator calls model.evaluate on
