"""Prints the L2 roofline probe (cdae_probe_l2) for a K x I table: row loads, vector reductions, both, and the
same reductions issued as bulk asynchronous reduces (cp.reduce.async.bulk) without / with the loads."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cdae_b200 import CDAE, CDAEConfig

def main():
    out = {}
    for K, I, visits in ((50, 50_000, 1_414_391), (100, 200_000, 1_414_391), (200, 27_000, 1_000_000)):
        rp = np.arange(0, 11, dtype=np.int64) * 2
        col = np.tile(np.array([0, 1], np.int32), 10)
        m = CDAE(CDAEConfig(loss="CE", num_dim=K)).reset(10, I, rp, col)
        r = {}
        for mode, name in ((1, "load"), (2, "red.v4"), (3, "load+red.v4"), (4, "bulk_red"), (5, "load+bulk_red")):
            g, ms = m.probe_l2(I, mode, visits)
            r[name] = {"GB/s_padded_rows": round(g, 1), "ms": round(ms, 4)}
        out["K=%d I=%d" % (K, I)] = r
        m.close()
    print(json.dumps(out, indent=1))

main()
