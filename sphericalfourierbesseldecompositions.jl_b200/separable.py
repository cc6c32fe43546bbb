"""SeparableArray: lazy outer product win[i, p] = phi[i] * mask[p]  (src/SeparableArrays.jl:53-122).

Only what the window path needs: the two named factors, shape, dense materialisation and indexing."""
from __future__ import annotations

import numpy as np


class SeparableArray:
    def __init__(self, phi, mask, name1="phi", name2="mask"):
        self.arr1 = np.ascontiguousarray(phi, dtype=np.float64)
        self.arr2 = np.ascontiguousarray(mask) if np.iscomplexobj(mask) else np.ascontiguousarray(mask, dtype=np.float64)
        if self.arr1.ndim != 1 or self.arr2.ndim != 1:
            raise ValueError("SeparableArray factors must be vectors")
        self.name1, self.name2 = name1, name2

    def __getattr__(self, name):
        if name in ("name1", "name2", "arr1", "arr2"):
            raise AttributeError(name)
        if name == self.name1:
            return self.arr1
        if name == self.name2:
            return self.arr2
        raise AttributeError(name)

    @property
    def shape(self):
        return (self.arr1.size, self.arr2.size)

    def __getitem__(self, idx):
        i, j = idx
        return np.multiply.outer(self.arr1[i], self.arr2[j])

    def dense(self):
        """win[:, :] of the reference: materialise the outer product."""
        return np.multiply.outer(self.arr1, self.arr2)

    def mean(self, axis=None):
        if axis is None:
            return self.arr1.mean() * self.arr2.mean()
        return self.arr2 * self.arr1.mean() if axis == 0 else self.arr1 * self.arr2.mean()
