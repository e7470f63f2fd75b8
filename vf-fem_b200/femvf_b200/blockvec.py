"""
Minimal ``BlockVector`` with the members the hot path touches.

The reference uses the author's un-vendored ``blockarray`` package for every state,
control and property vector (``/root/reference/src/femvf/models/transient.py:254-263,
711-718, 782-795, 809-814, 920``; ``statefile.py:97-105, 231-259, 281-291, 371-379``).
That package is not installable here, so its contract is restated from those call
sites: labelled 1D collection of numpy arrays, dict- and
slice-style access returning views, blockwise assignment and arithmetic.
"""

from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

_PARALLEL_COPY_MIN = 1 << 21   # elements
_POOL = None


def _copy_pool():
    """(pool, workers): threads for large host-to-host copies.  One process per GPU shares the
    host's cores with its sibling ranks (torchrun sets LOCAL_WORLD_SIZE), so the pool is sized
    to this rank's share, between 2 and 8 threads."""
    global _POOL
    import os
    if _POOL is not None and _POOL[2] != os.getpid():
        _POOL = None   # forked child: the parent's worker threads do not exist here
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        try:
            ranks = max(int(os.environ.get('LOCAL_WORLD_SIZE', '1')), 1)
        except ValueError:
            ranks = 1
        workers = min(max((os.cpu_count() or 2) // ranks, 2), 8)
        _POOL = (ThreadPoolExecutor(max_workers=workers), workers, os.getpid())
    return _POOL[:2]


def _parallel_copy(pairs):
    """Copy every (target, value) pair; large 1-D blocks are split into cache-line aligned chunks
    so that all threads of the pool move data (numpy releases the GIL inside the copy; a single
    thread does ~10 GB/s, far below what the page-locked upload that follows can take)."""
    pool, workers = _copy_pool()
    total = sum(t.size for t, _ in pairs)
    jobs = []
    for t, v in pairs:
        v = np.reshape(v, t.shape)
        # about two chunks per thread over all blocks, shared out by size
        per_block = int(round(2.0 * workers * t.size / max(total, 1)))
        if per_block > 1 and t.ndim == 1 and t.size >= (1 << 18):
            bounds = (np.linspace(0, t.size, per_block + 1).astype(np.int64) // 8) * 8
            bounds[-1] = t.size
            jobs += [(t[a:b], v[a:b]) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        else:
            jobs.append((t, v))
    list(pool.map(lambda tv: np.copyto(tv[0], tv[1]), jobs))


def _as_block(x):
    if isinstance(x, np.ndarray):
        return x
    return np.atleast_1d(np.asarray(x, dtype=np.float64))


class _SubIndexer:
    def __init__(self, bvec: 'BlockVector'):
        self._b = bvec

    def __getitem__(self, key):
        return self._b._get(key)

    def __setitem__(self, key, value):
        self._b[key] = value


class BlockVector:
    def __init__(self, vecs: Sequence, shape=None, labels=None):
        self._vecs = tuple(_as_block(v) for v in vecs)
        if labels is None:
            labels = (tuple(str(i) for i in range(len(self._vecs))),)
        self._labels = (tuple(labels[0]),)
        if len(self._labels[0]) != len(self._vecs):
            raise ValueError("number of labels must match number of blocks")

    # --- structure ------------------------------------------------------------
    @property
    def vecs(self):
        return self._vecs

    blocks = vecs
    sub_blocks = vecs

    @property
    def labels(self):
        return self._labels

    @property
    def shape(self):
        return (len(self._vecs),)

    @property
    def size(self):
        return len(self._vecs)

    @property
    def bshape(self):
        return (tuple(v.size for v in self._vecs),)

    @property
    def mshape(self):
        return (sum(v.size for v in self._vecs),)

    @property
    def sub(self):
        return _SubIndexer(self)

    def keys(self):
        return list(self._labels[0])

    def items(self):
        return list(zip(self._labels[0], self._vecs))

    sub_items = items

    def __contains__(self, key):
        return key in self._labels[0]

    def __iter__(self):
        return iter(self._vecs)

    def __len__(self):
        return len(self._vecs)

    # --- access ---------------------------------------------------------------
    def _index_of(self, key):
        if isinstance(key, str):
            try:
                return self._labels[0].index(key)
            except ValueError:
                raise KeyError(key) from None
        return int(key)

    def _get(self, key):
        if isinstance(key, (str, int, np.integer)):
            return self._vecs[self._index_of(key)]
        if isinstance(key, slice):
            idx = range(*key.indices(len(self._vecs)))
        elif isinstance(key, (list, tuple)):
            idx = [self._index_of(k) for k in key]
        else:
            raise TypeError(f"invalid BlockVector index {key!r}")
        return BlockVector([self._vecs[i] for i in idx],
                           labels=(tuple(self._labels[0][i] for i in idx),))

    def __getitem__(self, key):
        return self._get(key)

    def __setitem__(self, key, value):
        target = self._get(key)
        if isinstance(target, BlockVector):
            if isinstance(value, BlockVector):
                if len(value) != len(target):
                    raise ValueError("block count mismatch in assignment")
                pairs = list(zip(target.vecs, value.vecs))
                if sum(t.size for t, _ in pairs) >= _PARALLEL_COPY_MIN:
                    _parallel_copy(pairs)
                else:
                    for t, v in pairs:
                        t[...] = np.reshape(v, t.shape)
            elif np.isscalar(value):
                for t in target.vecs:
                    t[...] = value
            else:
                value = np.asarray(value, dtype=np.float64).reshape(-1)
                if value.size != target.mshape[0]:
                    raise ValueError("size mismatch in assignment")
                off = 0
                for t in target.vecs:
                    t[...] = value[off:off + t.size].reshape(t.shape)
                    off += t.size
        else:
            target[...] = value

    def set_mono(self, value):
        self[:] = value

    def to_mono_ndarray(self):
        return np.concatenate([np.ravel(v) for v in self._vecs]) if self._vecs else np.zeros(0)

    def copy(self):
        return BlockVector([np.array(v, copy=True) for v in self._vecs], labels=self._labels)

    # --- arithmetic -----------------------------------------------------------
    def _binary(self, other, op):
        if isinstance(other, BlockVector):
            return BlockVector([op(a, b) for a, b in zip(self._vecs, other.vecs)],
                               labels=self._labels)
        return BlockVector([op(a, other) for a in self._vecs], labels=self._labels)

    def __add__(self, o):
        return self._binary(o, np.add)

    __radd__ = __add__

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __rsub__(self, o):
        return (-self).__add__(o)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._binary(o, np.divide)

    def __neg__(self):
        return BlockVector([-a for a in self._vecs], labels=self._labels)

    def norm(self):
        return float(np.sqrt(sum(float(np.vdot(v, v)) for v in self._vecs)))

    def __repr__(self):
        return f"BlockVector(labels={self._labels}, bshape={self.bshape})"


def concatenate(bvecs: Iterable[BlockVector], labels=None) -> BlockVector:
    vecs, lab = [], []
    for b in bvecs:
        vecs += list(b.vecs)
        lab += list(b.labels[0])
    if labels is not None:
        lab = list(labels[0])
    return BlockVector(vecs, labels=(tuple(lab),))


def chunk(bvec: BlockVector, sizes: Sequence[int]):
    out, off = [], 0
    for n in sizes:
        out.append(bvec[off:off + n])
        off += n
    return tuple(out)


def norm(bvec: BlockVector) -> float:
    return bvec.norm()


def dot(a: BlockVector, b: BlockVector) -> float:
    return float(sum(float(np.vdot(x, y)) for x, y in zip(a.vecs, b.vecs)))
