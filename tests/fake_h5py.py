"""A minimal in-repo stand-in for h5py (absent from this image), enough for hpc.checkpoint: File as a context
manager, create_dataset (data= or shape=/dtype=), slice assignment and reads, attrs, groups, keys(), `in`.
Files are persisted as pickles so that a second File(path, 'r') sees what the first wrote.  TEST INFRASTRUCTURE."""
import pickle
from pathlib import Path

import numpy as np


class _Dataset:
    def __init__(self, data=None, shape=None, dtype=None, **kw):
        self.kw = kw
        if data is not None:
            self.arr = np.array(data) if dtype is None else np.array(data, dtype=dtype)
        else:
            self.arr = np.zeros(shape, dtype=dtype)
        self.writes = 0

    def __setitem__(self, key, value):
        self.arr[key] = value
        self.writes += 1

    def __getitem__(self, key):
        return self.arr[key].copy() if isinstance(self.arr[key], np.ndarray) else self.arr[key]

    @property
    def shape(self):
        return self.arr.shape

    @property
    def dtype(self):
        return self.arr.dtype


class _Group:
    def __init__(self):
        self.items, self.attrs = {}, {}

    def create_dataset(self, name, data=None, shape=None, dtype=None, **kw):
        ds = _Dataset(data=data, shape=shape, dtype=dtype, **kw)
        self.items[name] = ds
        return ds

    def create_group(self, name):
        g = _Group()
        self.items[name] = g
        return g

    def keys(self):
        return list(self.items.keys())

    def __getitem__(self, name):
        return self.items[name]

    def __contains__(self, name):
        return name in self.items


class File(_Group):
    def __init__(self, path, mode='r'):
        super().__init__()
        self.path, self.mode = Path(path), mode
        if mode == 'r':
            with open(self.path, 'rb') as fh:
                self.items, self.attrs = pickle.load(fh)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.mode != 'r':
            with open(self.path, 'wb') as fh:
                pickle.dump((self.items, self.attrs), fh)
        return False
