"""N2 measurement: "user item" text -> training structure, native loader (cdae_dataset_load_pairs:
text -> CSR) vs the reference (Data::load -> vector<Instance>, then RecsysModelBase::reset -> hash of
hashes; verbatim headers through oracle/_ref, which also allocates the model's parameters).
CPU only.  usage: python tests/experiments/loader_bench.py [users items mean]"""
import ctypes as C
import os
import resource
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cdae_b200 import Dataset, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    a = sys.argv[1:]
    U, I, mean = (int(a[0]), int(a[1]), float(a[2])) if len(a) >= 3 else (100_000, 50_000, 30.0)
    d = synth.make_dataset(U, I, mean_train=mean, seed=1)
    rp, col = d["train_row_ptr"], d["train_col"]
    path = os.path.join(tempfile.mkdtemp(), "pairs.txt")
    orc.write_pairs(path, rp, col)
    size = os.path.getsize(path)
    t = time.perf_counter()
    ds = Dataset(path)
    t_load = time.perf_counter() - t
    t = time.perf_counter()
    ds.random_split_by_feature_group(0.2, seed=1)
    t_split = time.perf_counter() - t
    n = ds.num_instances
    print("file %.1f MB, %d pairs, %d users, %d items" % (size / 1e6, n, ds.num_users, ds.num_items))
    print("native : load+CSR %.3f s (%.1f M pairs/s, %.0f MB/s), split+2 CSRs %.3f s" % (t_load, n / t_load / 1e6, size / t_load / 1e6, t_split))
    if orc.have_reference():
        L = orc._ref()
        r0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        t = time.perf_counter()
        h = L.ref_create((C.c_double * 4)(0.01, 0.1, 0.5, 0.0), (C.c_int32 * 11)(5, 4, 1, 1, 1, 0, 1, 0, 1, 0, 0), path.encode())
        t_ref = time.perf_counter() - t
        r1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        assert (L.ref_num_users(h), L.ref_num_items(h)) == (ds.num_users, ds.num_items)
        L.ref_destroy(h)
        print("reference: Data::load + CDAE::reset (K=4) %.3f s (%.2f M pairs/s), +%.0f MB resident" % (t_ref, n / t_ref / 1e6, (r1 - r0) / 1024.0))
        print("ratio %.1fx" % (t_ref / t_load))


if __name__ == "__main__":
    main()
