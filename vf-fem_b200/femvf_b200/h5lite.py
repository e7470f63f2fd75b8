"""
Tiny stand-in for the subset of ``h5py`` that ``StateFile`` uses.

h5py/libhdf5 are not installable in this environment.  ``File`` offers groups, resizable
datasets with numpy indexing, ``require_group/require_dataset/create_dataset`` and
persistence to a single ``.npz`` archive whose keys are the HDF5 dataset paths, so the
reference's layout (``/root/reference/src/femvf/statefile.py:163-270``) is preserved key for
key.  When h5py is importable ``statefile`` uses it instead.
"""

from __future__ import annotations

import os

import numpy as np


class Dataset:
    def __init__(self, name, shape, dtype, data=None, maxshape=None, chunks=None):
        self.name = name
        self.dtype = np.dtype(dtype)
        # storage grows geometrically along axis 0 (StateFile appends one row per step);
        # _n is the logical length, _buf the allocated block
        self._buf = np.zeros(shape, dtype=self.dtype)
        self._n = self._buf.shape[0] if self._buf.ndim else 0
        if data is not None:
            self._buf[...] = data
        self.maxshape = maxshape
        self.chunks = chunks

    @property
    def _data(self):
        return self._buf[:self._n] if self._buf.ndim else self._buf

    @property
    def shape(self):
        return self._data.shape

    @property
    def size(self):
        return self._data.size

    def resize(self, size, axis=None):
        if axis is None:
            new_shape = tuple(size)
        else:
            new_shape = list(self._data.shape)
            new_shape[axis] = size
            new_shape = tuple(new_shape)
        if self._buf.ndim and new_shape[1:] == self._buf.shape[1:]:
            n_new = new_shape[0]
            if n_new > self._buf.shape[0]:
                cap = max(n_new, 2 * self._buf.shape[0], 16)
                buf = np.zeros((cap,) + new_shape[1:], dtype=self.dtype)
                buf[:self._n] = self._buf[:self._n]
                self._buf = buf
            elif n_new < self._n:
                self._buf[n_new:self._n] = 0
            self._n = n_new
            return
        new = np.zeros(new_shape, dtype=self.dtype)
        sl = tuple(slice(0, min(a, b)) for a, b in zip(self._data.shape, new_shape))
        new[sl] = self._data[sl]
        self._buf = new
        self._n = new.shape[0] if new.ndim else 0

    def __getitem__(self, key):
        return self._data[key]

    def __setitem__(self, key, value):
        self._data[key] = value

    def __len__(self):
        return self._data.shape[0]


class Group:
    def __init__(self, name='/'):
        self.name = name
        self._items = {}

    def _split(self, path):
        return [p for p in path.split('/') if p]

    def _walk(self, path, create=False):
        node = self
        parts = self._split(path)
        for p in parts[:-1]:
            if p not in node._items:
                if not create:
                    raise KeyError(path)
                node._items[p] = Group(node.name.rstrip('/') + '/' + p)
            node = node._items[p]
        return node, (parts[-1] if parts else None)

    def __contains__(self, path):
        try:
            node, leaf = self._walk(path)
        except KeyError:
            return False
        return leaf in node._items

    def __getitem__(self, path):
        node, leaf = self._walk(path)
        return node._items[leaf]

    def __setitem__(self, path, value):
        node, leaf = self._walk(path, create=True)
        value = np.asarray(value)
        node._items[leaf] = Dataset(leaf, value.shape, value.dtype, data=value)

    def keys(self):
        return self._items.keys()

    def items(self):
        return self._items.items()

    def require_group(self, path):
        node, leaf = self._walk(path, create=True)
        if leaf not in node._items:
            node._items[leaf] = Group(node.name.rstrip('/') + '/' + leaf)
        return node._items[leaf]

    def create_dataset(self, path, shape=None, dtype=np.float64, data=None, maxshape=None,
                       chunks=None, **kwargs):
        node, leaf = self._walk(path, create=True)
        if shape is None:
            shape = np.shape(data)
        node._items[leaf] = Dataset(leaf, shape, dtype, data=data, maxshape=maxshape,
                                    chunks=chunks)
        return node._items[leaf]

    def require_dataset(self, path, shape, dtype=np.float64, data=None, exact=False, **kwargs):
        if path in self:
            return self[path]
        return self.create_dataset(path, shape=shape, dtype=dtype, data=data, **kwargs)

    def _flatten(self, prefix=''):
        out = {}
        for key, item in self._items.items():
            p = f'{prefix}/{key}' if prefix else key
            if isinstance(item, Group):
                out.update(item._flatten(p))
            else:
                out[p] = item._data
        return out


class File(Group):
    def __init__(self, fname, mode='r', **kwargs):
        super().__init__('/')
        self.filename = fname
        self.mode = mode
        if mode in ('r', 'a', 'r+') and os.path.exists(self._path()):
            with np.load(self._path(), allow_pickle=False) as z:
                for key in z.files:
                    self[key] = z[key]
        elif mode == 'r':
            raise FileNotFoundError(fname)

    def _path(self):
        return self.filename if self.filename.endswith('.npz') else self.filename + '.npz'

    def flush(self):
        if self.mode != 'r':
            np.savez(self._path(), **self._flatten())

    def close(self):
        self.flush()
