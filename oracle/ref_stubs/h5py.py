"""Import stub for h5py (not installed here).  TEST INFRASTRUCTURE, authoring container only.

The reference's core/dataset.py reads its training set through `h5py.File(path, 'r')[key][index]`.  So that the
UNMODIFIED dataset code can be run against a synthetic training set (oracle/gen_golden_feed.py,
tests/test_feed.py), `File` here opens an `.npz` archive with the same keys and hands out numpy arrays: the subset of
the h5py API the reference touches (`keys`, `in`, `[]`, `len`, `close`, context manager; `swmr=` is accepted).  One
h5py rule is enforced because the reference relies on it: fancy indices must be increasing (dataset.py:346 "hdf5 takes
increasing index order")."""
import numpy as np


class _Dataset:
    def __init__(self, a):
        self._a = a

    def __getitem__(self, idx):
        for i in (idx if isinstance(idx, tuple) else (idx,)):
            if isinstance(i, (list, np.ndarray)) and np.ndim(i) == 1 and len(i) > 1:
                if not (np.diff(np.asarray(i)) > 0).all():
                    raise TypeError("Indexing elements must be in increasing order")
        return self._a[idx]

    def __len__(self):
        return len(self._a)

    @property
    def shape(self):
        return self._a.shape

    @property
    def dtype(self):
        return self._a.dtype


class File:
    def __init__(self, path, mode="r", swmr=False, **kw):
        if mode != "r":
            raise NotImplementedError("the stub only reads")
        with np.load(path) as z:
            self._d = {k: z[k] for k in z.files}

    def keys(self):
        return self._d.keys()

    def __contains__(self, k):
        return k in self._d

    def __getitem__(self, k):
        return _Dataset(self._d[k])

    def __len__(self):
        return len(self._d)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
