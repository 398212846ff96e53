"""Python mirror of the data path in front of the hot path (include/cdae_b200.h, cdae_dataset_*):
`Data::load(file, RECSYS, ...)` + `random_split_by_feature_group` of the reference
(src/base/data-inl.hpp:45-64, 231-272) straight to the CSR `CDAE.reset` takes."""
import ctypes as C

import numpy as np

from . import _lib


class Dataset:
    """`Data` for "user item" text files: dense ids in first-seen order (instance-inl.hpp:22-37)."""

    def __init__(self, path, delimiters=" ", skip_header=False, _cache=False):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        if _cache:
            _lib.check(self._L.cdae_dataset_load(str(path).encode(), C.byref(self._h)))
        else:
            _lib.check(self._L.cdae_dataset_load_pairs(str(path).encode(), delimiters.encode(), int(skip_header),
                                                       C.byref(self._h)))
        u, i, n = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self._L.cdae_dataset_info(self._h, C.byref(u), C.byref(i), C.byref(n)))
        self.num_users, self.num_items, self.num_instances = u.value, i.value, n.value

    def save(self, path):
        """Data::save (data.hpp:25-33) in the library's own cache format: ids, instances and the split if made."""
        _lib.check(self._L.cdae_dataset_save(self._h, str(path).encode()))

    @classmethod
    def load(cls, path):
        """Data::load(cache file) (data.hpp:52-60): a data set written by save()."""
        return cls(path, _cache=True)

    def close(self):
        if self._h:
            self._L.cdae_dataset_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def random_split_by_feature_group(self, test_ratio=0.2, seed=20141119):
        """Data::random_split_by_feature_group(train, test, 0, test_ratio); returns (train, test) CSR pairs."""
        _lib.check(self._L.cdae_dataset_split(self._h, float(test_ratio), int(seed)))
        return self.csr("train"), self.csr("test")

    def csr(self, which="all"):
        w = {"all": 0, "train": 1, "test": 2}[which]
        nnz = C.c_int64()
        _lib.check(self._L.cdae_dataset_nnz(self._h, w, C.byref(nnz)))
        rp = np.zeros(self.num_users + 1, np.int64)
        col = np.zeros(max(nnz.value, 1), np.int32)
        _lib.check(self._L.cdae_dataset_csr(self._h, w, rp.ctypes.data_as(_lib.i64p), col.ctypes.data_as(_lib.i32p)))
        return rp, col[:nnz.value]

    def raw_id(self, group, idx):
        s = C.c_char_p()
        _lib.check(self._L.cdae_dataset_raw_id(self._h, int(group), int(idx), C.byref(s)))
        return s.value.decode()
