"""the other models apps/yelp/yelp.cpp instantiates (Popularity — always; ItemCF, IMF,
BPR on request) exist in the drop-in host tree (cdae_b200/host/model/recsys/) with the reference's class
surface, and the reference's UNCHANGED app built on them reproduces the reference binary's evaluation
tables on the same train / test split: Popularity and IMF / BPR exactly (they consume rand() in the same
order as the reference, and yelp.cpp never calls srand), ItemCF up to the order of exactly tied scores.
CPU only (these models do not touch the CUDA engine)."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from tests.test_host_app import B200, REF, HOST, _write_pairs

pytestmark = pytest.mark.skipif(not (os.path.exists(B200) and os.path.exists(REF)),
                                reason="app binaries are built where /root/reference exists (cdae_b200/host/Makefile)")


def _tables(log_text):
    """All Solver tables of a log: list of lists of rows [iter, time, loss, P@1 .. MAP@10, test time]."""
    tables = []
    for line in log_text.splitlines():
        m = re.search(r"solver-inl\.hpp:\d+\]\s+(\d+)\|(.*)$", line)
        if m:
            vals = [float(x) for x in m.group(2).strip().strip("|").split("|")]
            if int(m.group(1)) == 0:
                tables.append([])
            tables[-1].append([int(m.group(1))] + vals)
    return tables


def test_host_headers_keep_the_reference_surface():
    d = os.path.join(HOST, "model", "recsys")
    want = {"popularity.hpp": ["_LIBCF_POPULARITY_HPP_", "class Popularity : public RecsysModelBase", "void reset(const Data& data_set)"],
            "similarity_base.hpp": ["_LIBCF_SIMILARITY_BASE_HPP_", "enum SimilarityType { Jaccard, Cosine }", "get_neighbors() const"],
            "itemcf.hpp": ["_LIBCF_ITEMCF_HPP_", "ItemCF(SimilarityType sim_type = Jaccard, size_t topk = 50)"],
            "imf.hpp": ["_LIBCF_IMF_HPP_", "struct IMFConfig", "virtual void train_one_instance(size_t uid, size_t iid, double rui)",
                        "double predict_user_item_rating(size_t uid, size_t iid) const", "DMatrix get_user_vecs()"],
            "bpr.hpp": ["_LIBCF_BPR_HPP_", "struct BPRConfig", "class BPR : public IMF",
                        "virtual void train_one_pair(size_t uid, size_t iid, size_t jid, double rui)"]}
    for f, needles in want.items():
        src = open(os.path.join(d, f)).read()
        for n in needles:
            assert n in src, (f, n)


def test_app_on_host_models_matches_the_reference_binary(tmp_path):
    base = tmp_path / "data"
    base.mkdir()
    (base / "log").mkdir()
    _write_pairs(str(base / "yelp_10core.txt"), U=700, I=500, mean=12.0, seed=9)
    for task in ("prepare", "split"):                      # ONE split, made by the reference binary, used by both
        subprocess.run([REF, "--task=" + task], cwd=str(base), capture_output=True, timeout=300)
    runs = {}
    for name, binary in (("ref", REF), ("ours", B200)):
        d = tmp_path / name
        shutil.copytree(str(base), str(d))
        runs[name] = {}
        for method in ("ITEMCF", "MF", "BPR"):
            r = subprocess.run([binary, "--task=test", "--method=" + method, "--num_dim=16", "--loss_type=LOG"],
                               cwd=str(d), capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, (name, method, r.stderr[-1500:])
            runs[name][method] = _tables(open(str(d / "log" / "yelp_implicit.log")).read())[-2:]   # the log file is appended to
    for method in ("ITEMCF", "MF", "BPR"):
        a, b = runs["ref"][method], runs["ours"][method]
        assert len(a) == len(b) == 2                       # Popularity (always, yelp.cpp:109-113) + the method
        pop_a, pop_b = np.array(a[0][0][3:11]), np.array(b[0][0][3:11])
        assert np.array_equal(pop_a, pop_b), (pop_a, pop_b)            # Popularity: identical metrics
        fa, fb = np.array(a[1][-1][3:11]), np.array(b[1][-1][3:11])
        if method == "ITEMCF":
            assert np.abs(fa - fb).max() <= 5e-3, (fa, fb)              # exact ties broken differently
        else:
            assert len(a[1]) == len(b[1]) == 51                          # iteration 0 + 50 epochs
            assert np.array_equal(fa, fb), (method, fa, fb)             # same rand() stream, same arithmetic
