"""Generates tests/golden/pairs_small.{txt,npz}: a "user item" text file with string ids, duplicate
pairs and empty lines, and the users' item sets as THE REFERENCE'S OWN LOADER sees them
(Data::load + RecsysModelBase::reset through oracle/_ref, verbatim reference headers).
Run in the dev container (needs /root/reference):  python tests/golden/make_pairs_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    users = ["u%03d" % x for x in rng.permutation(40)]
    items = ["B00%04X" % x for x in rng.permutation(90)]
    lines = []
    for u in users:
        for it in rng.choice(items, size=int(rng.integers(3, 15)), replace=False):
            lines.append("%s %s" % (u, it))
    rng.shuffle(lines)
    lines = lines + lines[:17]                      # duplicate pairs
    lines.insert(5, "")                             # empty lines are skipped
    lines.insert(40, "")
    path = os.path.join(HERE, "pairs_small.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    orc.build(ref=True)
    L = orc._ref()
    import ctypes as C
    cfg_d = (C.c_double * 4)(0.01, 0.1, 0.5, 0.0)
    cfg_i = (C.c_int32 * 11)(5, 4, 1, 1, 1, 0, 1, 0, 1, 0, 0)
    h = L.ref_create(cfg_d, cfg_i, path.encode())
    U, I = L.ref_num_users(h), L.ref_num_items(h)
    rows = []
    buf = np.zeros(I, np.int64)
    for u in range(U):
        n = L.ref_user_items(h, u, buf.ctypes.data_as(orc.i64p), I)
        rows.append(np.sort(buf[:n].copy()))
    L.ref_destroy(h)
    rp = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    np.savez(os.path.join(HERE, "pairs_small.npz"), U=U, I=I, row_ptr=rp, col=np.concatenate(rows).astype(np.int32))
    print("users", U, "items", I, "pairs", rp[-1])


if __name__ == "__main__":
    main()
