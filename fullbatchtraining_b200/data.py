"""Minimal data plumbing for the hot path: synthetic CIFAR-shaped tensors and a block loader over (pinned) host
tensors with the reference DataLoader protocol that ``train`` consumes (iteration yields (inputs fp32 NCHW, labels
int64) blocks of ``batch_size``, ``drop_last=True``, ``len()`` = number of blocks; reference
fullbatch/data/data_preparation.py:56-72), plus a reader of the reference's LMDB record format into the resident
uint8 dataset (``load_lmdb_records``).  Downloading / writing datasets is out of scope."""
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


def load_lmdb_records(path_or_env, device="cpu", limit=None):
    """Read a database written by the reference's ``LMDBDataset`` (fullbatch/data/lmdb_datasets.py:132-162 read path,
    :228-271 write path) into the resident format of the B200 path: uint8 images [N,32,32,3] (HWC) + int64 labels.

    On-disk layout: records keyed by the ASCII decimal index holding the raw uint8 pixels of one image, either CHW
    (``ToTensor`` first in the live transform) or HWC, plus pickled ``__len__``, ``__keys__``, ``__labels__`` and
    ``__shape__`` entries (lmdb_datasets.py:70-75).  With ``rounds`` > 1 the database holds N x pre-augmented copies
    (the "10x / 40x CIFAR" sets), which simply become more resident images.

    `path_or_env`: a path (opened read-only through the ``lmdb`` package, which must then be installed) or an object
    with lmdb's ``begin(write=False)`` -> transaction (``get(key)``) protocol.
    """
    import pickle

    env = path_or_env
    if isinstance(path_or_env, (str, bytes)):
        try:
            import lmdb
        except ImportError as e:  # not part of this image: the reader is exercised against an API stub in the tests
            raise RuntimeError("reading an LMDB file needs the `lmdb` package (pip install lmdb)") from e
        env = lmdb.open(path_or_env, subdir=False, readonly=True, lock=False, readahead=True, meminit=False)
    with env.begin(write=False) as txn:
        meta = {k: txn.get(k) for k in (b"__len__", b"__keys__", b"__labels__", b"__shape__")}
        if any(v is None for v in meta.values()):
            raise ValueError("LMDB database is unfinished or damaged (missing __len__/__keys__/__labels__/__shape__)")
        length = pickle.loads(meta[b"__len__"])
        keys = pickle.loads(meta[b"__keys__"])
        labels = pickle.loads(meta[b"__labels__"])
        shape = tuple(pickle.loads(meta[b"__shape__"]))
        n = length if limit is None else min(limit, length)
        if shape == (3, 32, 32):
            chw = True
        elif shape == (32, 32, 3):
            chw = False
        else:
            raise ValueError(f"unsupported record shape {shape}: the path handles 32x32 RGB images")
        out = torch.empty(n, 32, 32, 3, dtype=torch.uint8)
        for i in range(n):
            buf = txn.get(keys[i])
            if buf is None or len(buf) != 3072:
                raise ValueError(f"record {i} is missing or has {0 if buf is None else len(buf)} bytes instead of 3072")
            img = torch.frombuffer(bytearray(buf), dtype=torch.uint8).view(shape)
            out[i] = img.permute(1, 2, 0) if chw else img
    y = torch.as_tensor(labels[:n], dtype=torch.int64)
    return out.to(device), y.to(device)
