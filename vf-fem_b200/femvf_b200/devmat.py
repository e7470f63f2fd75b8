"""
Device-resident CSR block returned by ``FenicsModel.assem_dres_dstate1`` for ``dF_u/du1``.

The reference hands back a PETSc matrix (``models/transient.py:384-406``) that the caller almost
always passes straight to ``solve_dres_dstate1`` (``transient.py:470-491``) or multiplies with a
vector.  Here the values stay where the assembly kernel wrote them (the engine's ``J`` array):

* ``A @ x`` / ``A.dot(x)`` run the device SpMV kernel (x is uploaded, y downloaded);
* ``solve_dres_dstate1`` recognises the object and solves with the resident values;
* anything that needs the numbers on the host -- ``.data``, ``.tocsr()``, ``.toarray()``,
  ``.diagonal()``, or any other scipy attribute -- downloads the values ONCE (0.45 GB at the
  benchmark size) and behaves like the ``scipy.sparse.csr_matrix`` it then wraps.

Copy-on-write keeps the reference's value semantics: the engine has a single ``J`` array, so the
model snapshots a still-referenced, not yet materialised matrix to a private device tensor before
the next assembly overwrites ``J`` (``FenicsModel._retire_live_jacobian``).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class DeviceCSR:
    __array_priority__ = 20.0   # numpy defers ``ndarray @ DeviceCSR`` to __rmatmul__

    def __init__(self, model, member: int, pinned: bool):
        self._model = model
        self._engine = model.engine
        self._member = member
        self._pinned = pinned
        self._snapshot = None      # device clone of the values once J has been reassembled
        self._host = None          # scipy csr_matrix once materialised
        N = self._engine.N
        self.shape = (N, N)
        self.dtype = np.dtype(np.float64)
        self.nnz = self._engine.nnz
        self.ndim = 2
        self.format = 'csr'

    # --- residency ----------------------------------------------------------------------------
    @property
    def on_device(self) -> bool:
        """True while the host copy has not been made."""
        return self._host is None

    def _is_live(self) -> bool:
        return self._snapshot is None and self._model._live_jacobian() is self

    def _detach(self):
        """Called by the model before ``J`` is overwritten: keep the values in a private tensor."""
        if self._host is None and self._snapshot is None:
            self._snapshot = self._engine.view('J', self._member).clone()

    def _values_tensor(self):
        return self._snapshot if self._snapshot is not None \
            else self._engine.view('J', self._member)

    def tocsr(self, copy: bool = False) -> sp.csr_matrix:
        if self._host is None:
            rowptr, colidx = self._model.csr_pattern()
            if self._snapshot is not None:
                vals = self._snapshot.cpu().numpy()
                self._snapshot = None
            else:
                vals = self._engine.download('J', self._member, pinned=self._pinned)
                if self._pinned:
                    vals = vals.copy()   # the staging buffer is reused by the next download
            self._host = sp.csr_matrix((vals, colidx, rowptr), shape=self.shape)
        return self._host.copy() if copy else self._host

    # --- products on the device ---------------------------------------------------------------
    def _matvec(self, x):
        import torch
        x = np.asarray(x, dtype=np.float64)
        if self._host is not None or not self._is_live() or x.ndim != 1:
            return self.tocsr() @ x
        e = self._engine
        xt = torch.as_tensor(np.ascontiguousarray(x), device=e.device)
        yt = torch.empty_like(xt)
        e.spmv(xt, yt, self._member)
        return yt.cpu().numpy()

    def dot(self, x):
        return self._matvec(x) if isinstance(x, np.ndarray) and x.ndim == 1 else self.tocsr().dot(x)

    def __matmul__(self, x):
        return self.dot(x)

    def __rmatmul__(self, x):
        return x @ self.tocsr()

    # --- scipy look-alike (host copy on first use) ---------------------------------------------
    @property
    def data(self):
        return self.tocsr().data

    @property
    def indices(self):
        return self._model.csr_pattern()[1]

    @property
    def indptr(self):
        return self._model.csr_pattern()[0]

    def toarray(self):
        return self.tocsr().toarray()

    def diagonal(self, k: int = 0):
        return self.tocsr().diagonal(k)

    def copy(self):
        return self.tocsr(copy=True)

    def __getattr__(self, name):
        # anything else scipy offers (tocsc, T, multiply, ...): materialise and delegate
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.tocsr(), name)

    def __add__(self, other):
        return self.tocsr() + other

    def __radd__(self, other):
        return other + self.tocsr()

    def __sub__(self, other):
        return self.tocsr() - other

    def __rsub__(self, other):
        return other - self.tocsr()

    def __mul__(self, other):
        return self.tocsr() * other

    def __rmul__(self, other):
        return other * self.tocsr()

    def __neg__(self):
        return -self.tocsr()

    def __repr__(self):
        where = 'device' if self.on_device else 'host'
        return f"<DeviceCSR {self.shape[0]}x{self.shape[1]}, {self.nnz} stored elements, on {where}>"
