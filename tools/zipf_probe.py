"""How much of the sampled decode's time is same-address reduction contention on the Zipf head?
Times decode_kernel (per 8,192-user minibatch, config B shape) on data sets of different item popularity skew."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from cdae_b200 import CDAE, CDAEConfig, synth

def main():
    c = bench.CONFIGS["B"]
    out = {}
    for alpha in (1.0, 0.5, 0.0):
        d = synth.make_dataset(c["users_per_gpu"], c["items"], c["mean"], seed=bench.SEED, alpha=alpha)
        m = CDAE(CDAEConfig(batch_users=8192, **bench.model_cfg(c))).reset(d["U"], d["I"], d["train_row_ptr"], d["train_col"])
        m.init_params(bench.SEED)
        for ep in range(3):
            m.train_one_iteration(seed=1, epoch=ep)
        m.profile(True)
        outs = 0
        for ep in range(3, 8):
            st = m.train_one_iteration(seed=1, epoch=ep)
            outs += st.outputs
        p = m.profile_get()
        cnt = np.bincount(d["train_col"], minlength=d["I"])
        out["alpha=%.1f" % alpha] = {"decode_us_per_launch": 1e3 * p["decode"][0] / p["decode"][1], "rows_per_launch": outs / p["decode"][1],
                                      "ns_per_krow": 1e6 * p["decode"][0] / outs * 1e3, "top_item_share_of_positives": float(cnt.max() / cnt.sum()),
                                      "scatter_us": 1e3 * p["scatter"][0] / p["scatter"][1], "gather_us": 1e3 * p["gather"][0] / p["gather"][1]}
        m.close()
    print(json.dumps(out, indent=1))
main()
