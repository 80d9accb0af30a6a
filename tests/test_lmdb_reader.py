"""Reader of the reference's LMDB record format (fullbatch/data/lmdb_datasets.py) against an lmdb-API stub: the `lmdb`
package is not part of this image, so the on-disk B+tree itself is not exercised."""
import pickle

import pytest
import torch

from fullbatchtraining_b200.data import load_lmdb_records


class _Txn:
    def __init__(self, store):
        self.store = store

    def get(self, key):
        return self.store.get(key)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _Env:
    """dict-backed stand-in with lmdb's begin() -> transaction -> get() protocol"""

    def __init__(self, store):
        self.store = store

    def begin(self, write=False):
        assert write is False
        return _Txn(self.store)


def _database(n, chw, seed=0):
    """what _create_database writes (lmdb_datasets.py:228-271)"""
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randint(0, 256, (n, 3, 32, 32), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 10, (n,), generator=g).tolist()
    store = {}
    for i in range(n):
        rec = imgs[i] if chw else imgs[i].permute(1, 2, 0)
        store[str(i).encode("ascii")] = rec.contiguous().numpy().tobytes()
    store[b"__keys__"] = pickle.dumps([str(i).encode("ascii") for i in range(n)])
    store[b"__labels__"] = pickle.dumps(labels)
    store[b"__len__"] = pickle.dumps(n)
    store[b"__shape__"] = pickle.dumps((3, 32, 32) if chw else (32, 32, 3))
    return store, imgs, labels


@pytest.mark.parametrize("chw", [True, False], ids=["CHW", "HWC"])
def test_records_become_the_resident_uint8_dataset(chw):
    store, imgs, labels = _database(37, chw)
    X, Y = load_lmdb_records(_Env(store))
    assert X.dtype == torch.uint8 and tuple(X.shape) == (37, 32, 32, 3) and Y.dtype == torch.int64
    assert torch.equal(X, imgs.permute(0, 2, 3, 1))
    assert Y.tolist() == labels
    X5, Y5 = load_lmdb_records(_Env(store), limit=5)
    assert torch.equal(X5, X[:5]) and torch.equal(Y5, Y[:5])


def test_damaged_databases_are_rejected():
    store, _, _ = _database(4, True)
    broken = dict(store)
    del broken[b"__labels__"]
    with pytest.raises(ValueError):
        load_lmdb_records(_Env(broken))
    short = dict(store)
    short[b"2"] = short[b"2"][:100]
    with pytest.raises(ValueError):
        load_lmdb_records(_Env(short))
    odd = dict(store)
    odd[b"__shape__"] = pickle.dumps((1, 28, 28))
    with pytest.raises(ValueError):
        load_lmdb_records(_Env(odd))


def test_path_without_the_lmdb_package_fails_loudly():
    try:
        import lmdb  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="lmdb"):
            load_lmdb_records("/nonexistent/train.lmdb")
