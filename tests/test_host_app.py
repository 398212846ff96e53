"""The reference's own app (apps/yelp/yelp.cpp, compiled UNCHANGED) on the drop-in
libcf::CDAE of cdae_b200/host/ — INTEGRATION.md.  The binaries are built where /root/reference
exists (cdae_b200/host/Makefile) and travel to the GPU box prebuilt."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "cdae_b200", "host")
B200 = os.path.join(HOST, "_build", "yelp_b200")
REF = os.path.join(HOST, "_build", "yelp_ref")
HAVE_REFERENCE = os.path.exists("/root/reference/apps/yelp/yelp.cpp")


def test_drop_in_header_keeps_the_reference_surface():
    """Same include guard, config fields and member names as the reference's cdae.hpp."""
    src = open(os.path.join(HOST, "model", "recsys", "cdae.hpp")).read()
    assert "#ifndef _LIBCF_CDAE_HPP_" in src and "class CDAE : public RecsysModelBase" in src
    for field in ("lambda", "learn_rate", "lt", "pt", "num_dim", "using_adagrad", "corruption_ratio",
                  "num_corruptions", "asymmetric", "user_factor", "linear", "num_neg", "scaled", "beta",
                  "linear_function", "tanh"):
        assert re.search(r"\b%s\s*=" % field, src), field
    for member in ("double data_loss(const Data& data_set, size_t sample_size = 0) const",
                   "double penalty_loss() const", "void reset(const Data& data_set)",
                   "void train_one_iteration(const Data& train_data)",
                   "DMatrix get_user_representations()", "virtual void pre_recommend()",
                   "std::vector<size_t> recommend(size_t uid, size_t topk,",
                   "void train_one_user_corruption(size_t uid,",
                   "get_corrputed_input(const std::unordered_map<size_t, double>& input_set,",
                   "DVector get_hidden_values(size_t uid, const std::unordered_map<size_t, double>& item_set,",
                   "double get_output_values(const DVector& z, size_t iid) const"):
        assert member in src, member
    assert "oracle" not in src.lower()          # the product never touches the checker


@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs /root/reference (dev container only)")
def test_unmodified_yelp_app_links_against_the_engine():
    from cdae_b200 import build
    build.build()
    subprocess.check_call(["make", "-s", "-C", HOST], timeout=600)
    assert os.path.exists(B200) and os.path.exists(REF)
    ldd = subprocess.run(["ldd", B200], capture_output=True, text=True).stdout
    assert "libcdae_b200.so" in ldd
    assert "libcdae_b200.so" not in subprocess.run(["ldd", REF], capture_output=True, text=True).stdout


def _write_pairs(path, U=1500, I=800, mean=14.0, seed=5):
    from cdae_b200 import synth
    d = synth.make_dataset(U, I, mean, seed=seed)
    with open(path, "w") as f:
        f.write("user item\n")
        for u in range(U):
            a = d["train_col"][d["train_row_ptr"][u]:d["train_row_ptr"][u + 1]]
            b = d["test_col"][d["test_row_ptr"][u]:d["test_row_ptr"][u + 1]]
            for i in np.concatenate([a, b]):
                f.write("u%d i%d\n" % (u, i))


def _table(log_text):
    """Rows of the Solver table (solver-inl.hpp:24-35,62-69): Iters|Time|Train Loss|P@1..MAP@10|TestTime;
    one table per Solver::train call — the last one is the CDAE run."""
    rows = []
    for line in log_text.splitlines():
        m = re.search(r"solver-inl\.hpp:\d+\]\s+(\d+)\|(.*)$", line)
        if m:
            vals = [float(x) for x in m.group(2).strip().strip("|").split("|")]
            if int(m.group(1)) == 0:
                rows = []
            rows.append([int(m.group(1))] + vals)
    return rows


def _run_app(binary, cwd, extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    os.makedirs(os.path.join(cwd, "log"), exist_ok=True)
    for args in (["--task=prepare"], ["--task=split"],
                 ["--task=test", "--method=CDAE", "--num_dim=20", "--loss_type=CE", "--cratio=0.5",
                  "--scaled=true", "--beta=1"]):
        r = subprocess.run([binary] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
        # yelp.cpp:88-104: `if (train) {..} if (test) {..} else return -1;` — prepare and split do their
        # work and then fall into that else, so only --task=test exits 0
        if args[0] == "--task=test":
            assert r.returncode == 0, (args, r.stderr[-2000:])
    for f in ("yelp.bin", "yelp.train.bin", "yelp.test.bin"):
        assert os.path.exists(os.path.join(cwd, f)), f
    return _table(open(os.path.join(cwd, "log", "yelp_implicit.log")).read())


@pytest.mark.gpu
def test_yelp_app_trains_on_the_gpu_like_the_reference(tmp_path):
    if not (os.path.exists(B200) and os.path.exists(REF)):
        pytest.skip("prebuilt app binaries missing (built only where /root/reference exists)")
    runs = {}
    for name, binary in (("b200", B200), ("ref", REF)):
        d = tmp_path / name
        d.mkdir()
        _write_pairs(str(d / "yelp_10core.txt"))
        runs[name] = _run_app(binary, str(d), {"CDAE_B200_BATCH_USERS": "64"})
    g, r = runs["b200"], runs["ref"]
    assert len(g) == 51 and len(r) == 51                       # iteration 0 + 50 epochs (yelp.cpp:197)
    # columns: iter, time, train loss, P@1, P@5, P@10, R@1, R@5, R@10, MAP@5, MAP@10, test time
    assert g[-1][2] < g[1][2]                                   # the training loss goes down
    map10_g = np.mean([row[10] for row in g[-10:]])
    map10_r = np.mean([row[10] for row in r[-10:]])
    map10_0 = g[0][10]
    print("MAP@10: untrained %.4f  b200 %.4f  reference %.4f" % (map10_0, map10_g, map10_r))
    assert map10_g > 3 * max(map10_0, 1e-3)                    # it learned
    # statistical parity (different RNG streams, minibatch vs online): same quality band
    assert abs(map10_g - map10_r) <= 0.25 * map10_r
    loss_g, loss_r = g[-1][2], r[-1][2]
    assert abs(loss_g - loss_r) <= 0.25 * abs(loss_r)
