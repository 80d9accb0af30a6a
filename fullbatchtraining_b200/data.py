"""Minimal data plumbing for the hot path: synthetic CIFAR-shaped tensors and a block loader over (pinned) host
tensors with the reference DataLoader protocol that ``train`` consumes (iteration yields (inputs fp32 NCHW, labels
int64) blocks of ``batch_size``, ``drop_last=True``, ``len()`` = number of blocks; reference
fullbatch/data/data_preparation.py:56-72).  The real CIFAR/LMDB pipeline of the reference is out of scope."""
import torch


def synthetic_cifar(n, seed=1234, device="cpu"):
    """randn images (zero mean / unit variance per channel, like normalised CIFAR-10) and uniform labels."""
    gen = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, 3, 32, 32, generator=gen, device=device)
    y = torch.randint(0, 10, (n,), generator=gen, device=device)
    return x, y


class HostBlockLoader:
    """Sequential blocks of a host dataset held in pinned memory (zero-copy slices, asynchronous H2D)."""

    def __init__(self, inputs, labels, batch_size, pin=True):
        assert inputs.shape[0] == labels.shape[0]
        self.inputs = inputs.contiguous().pin_memory() if pin and torch.cuda.is_available() else inputs.contiguous()
        self.labels = labels.contiguous().pin_memory() if pin and torch.cuda.is_available() else labels.contiguous()
        self.batch_size = min(batch_size, inputs.shape[0])
        self.sampler = self

    def set_epoch(self, *args, **kwargs):  # data_preparation.py:58-62
        pass

    def __len__(self):
        return self.inputs.shape[0] // self.batch_size  # drop_last

    def __iter__(self):
        b = self.batch_size
        for i in range(len(self)):
            yield self.inputs[i * b:(i + 1) * b], self.labels[i * b:(i + 1) * b]
