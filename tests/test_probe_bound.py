"""The argument behind the probe pass of cdae_topn_build (csrc/topn_tc.cuh "Probe pass"), checked on the CPU with
plain numpy: approximate scores within eps of the exact ones, start threshold thr0 = a_k - 2 eps - tiny from the
k-th best APPROXIMATE score of any subset of the unrated items.  Then (1) no item the sweep skips (approx <= thr0)
can be in the exact top-k, and (2) the re-rank's verification test thr + eps < (k-th best exact score) holds for
thr0.  The reference's list is the exact top-k of the unrated items (cdae.hpp:162-196), so a sweep that starts at
thr0 proposes every item of it."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def thr0_from_probe(approx_probe, k, eps):
    """probe_thr_kernel: k-th largest approximate score of the probe candidates, minus 2 eps and a tiny margin;
    -inf when the probe holds fewer than k unrated items."""
    if approx_probe.size < k:
        return -np.inf
    kth = np.sort(approx_probe)[::-1][k - 1]
    return np.float32(kth) - np.float32(2.0) * np.float32(eps) - np.float32(1e-6) * abs(np.float32(kth)) - np.float32(1e-30)


@settings(max_examples=300, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(20, 400), k=st.integers(1, 16), frac=st.floats(0.02, 1.0),
       eps=st.floats(1e-4, 0.5), ties=st.booleans())
def test_items_below_the_start_threshold_cannot_be_in_the_list(seed, n, k, frac, eps, ties):
    rng = np.random.default_rng(seed)
    exact = rng.normal(size=n).astype(np.float32)
    if ties:
        exact = np.round(exact * 4) / 4                      # many equal scores
    # (eps_u of pack_z_bf16_kernel is thousands of fp32 ulps of the scores — bf16 operand rounding — so the
    #  fp32 rounding of this construction stays inside the 1 % it leaves free)
    approx = (exact.astype(np.float64) + rng.uniform(-eps, eps, size=n) * 0.99).astype(np.float32)
    assert np.all(np.abs(approx.astype(np.float64) - exact) <= eps)
    m = max(1, int(frac * n))
    probe = rng.choice(n, size=m, replace=False)              # ANY subset of the unrated items
    thr0 = thr0_from_probe(approx[probe], k, eps)
    if n < k:
        return
    kth_exact = np.sort(exact)[::-1][k - 1]
    skipped = approx <= thr0
    # (1) a skipped item is strictly below the k-th best exact score: it is in no exact top-k list,
    #     whatever the tie rule
    assert np.all(exact[skipped] < kth_exact)
    # (2) the verification test of the re-rank holds at the start threshold
    if np.isfinite(thr0):
        assert float(thr0) + eps < float(kth_exact)
    # so the candidates (approx > thr0) contain the whole list
    top = np.argsort(-exact, kind="stable")[:k]
    assert not np.any(skipped[top])


def test_fewer_than_k_probe_items_gives_no_threshold():
    assert thr0_from_probe(np.array([1.0, 2.0], np.float32), 10, 0.1) == -np.inf


@pytest.mark.parametrize("I,knob,want", [(400, 1, 0), (600, 1, 256), (4095, 1, 256), (4096, 1, 512), (50000, 1, 1024),
                                         (50000, 2, 512), (50000, 16, 4096), (200000, 64, 8192), (6310, 8, 768)])
def test_probe_table_size_rule(I, knob, want):
    """tc_probe_thresholds (topn_api.inl): one 256-item tile per 2,048 items, at most 4 (knob n >= 2: at most
    min(n, 32)); no probe below 512 items.  Mirrors tests/test_gpu_topn_probe.py::probe_size."""
    from tests.test_gpu_topn_probe import probe_size
    assert probe_size(I, knob) == want
