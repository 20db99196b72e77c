"""Import-time stand-in for `webdataset` (not installed): the reference's worker only uses these three names
for an isinstance() dispatch in `_create_dataloader` (reference: tf_yarn/pytorch/tasks/worker.py:49-64)."""


class WebDataset:
    pass


class DataPipeline:
    pass


class WebLoader:
    def __init__(self, *a, **k):
        raise RuntimeError("webdataset is not available offline (bench/shims)")
