"""Framework-agnostic distributed run: the user function sets up ``torch.distributed`` itself.

(reference: tf_yarn/examples/pytorch/pytorch_distributed_example.py:38-138) -- the MNIST CNN
(1,199,882 parameters), SGD lr 1e-4, a global batch of 128 split over the ranks, resume from the
latest checkpoint, 2 workers x 2 processes.  ``tf_yarn_b200.distributed`` only provides rank, world
size and the master address (``get_task``).
"""
import logging
import os
import tempfile

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch.nn.parallel import DistributedDataParallel as DDP

from tf_yarn_b200.distributed import run_on_yarn
from tf_yarn_b200.distributed.task import get_task
from tf_yarn_b200.models.mnist_cnn import TorchMnistCnn
from tf_yarn_b200.pytorch import model_ckpt
from tf_yarn_b200.topologies import NodeLabel, TaskSpec

logging.basicConfig(level="INFO")
MODEL_DIR = os.path.join(tempfile.gettempdir(), "tf_yarn_b200_test", "pytorch_distributed")
BATCH_SIZE = 128
LABEL = NodeLabel.GPU if torch.cuda.is_available() else NodeLabel.CPU


def experiment_fn(local_rank: int) -> None:
    task = get_task(local_rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = task.master_address, str(task.master_port)
    use_cuda = task.device is not None and torch.cuda.is_available()
    device = torch.device(f"cuda:{task.device}" if use_cuda else "cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    dist.init_process_group("nccl" if use_cuda else "gloo", rank=task.task_id, world_size=task.world_size)
    g = torch.Generator().manual_seed(0)
    dataset = torch.utils.data.TensorDataset(torch.rand(4096, 1, 28, 28, generator=g),
                                             torch.randint(0, 10, (4096,), generator=g))
    sampler = torch.utils.data.distributed.DistributedSampler(dataset)
    loader = torch.utils.data.DataLoader(dataset, batch_size=BATCH_SIZE // task.world_size, sampler=sampler,
                                         drop_last=True)
    model = TorchMnistCnn().to(device)
    optimizer = torch.optim.SGD(model.parameters(), lr=1e-4)
    ckpt = model_ckpt.load_latest_ckpt(MODEL_DIR, model, optimizer, device)
    start_epoch = ckpt["epoch"] + 1 if ckpt else 0
    ddp_model = DDP(model, device_ids=[device.index] if use_cuda else None)
    for epoch in range(start_epoch, start_epoch + int(os.environ.get("EXAMPLE_EPOCHS", "1"))):
        sampler.set_epoch(epoch)
        for x, y in loader:
            optimizer.zero_grad()
            loss = F.nll_loss(ddp_model(x.to(device)), y.to(device))
            loss.backward()
            optimizer.step()
        if task.task_id == 0:
            model_ckpt.save_ckpt(MODEL_DIR, ddp_model, optimizer, epoch)
        dist.barrier()
    dist.destroy_process_group()


def main():
    return run_on_yarn(experiment_fn, task_specs={
        "worker": TaskSpec(memory="4 GiB", vcores=4, instances=2, nb_proc_per_worker=2, label=LABEL)})


if __name__ == "__main__":
    print(main())
