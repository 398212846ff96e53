"""The native data path (cdae_dataset_*: SURVEY.md §8f N2) against the reference's own loader.
Host-only code: runs without a GPU.

  * golden: tests/golden/pairs_small.txt -> the users' item sets exactly as Data::load +
    RecsysModelBase::reset of the VERBATIM reference produce them (tests/golden/make_pairs_golden.py)
  * live (dev container only): a larger random file through oracle/_ref
  * tokenizer rules of split_line (file_utils.hpp:15-25; known answer in test/file_test.hpp:14-23),
    line rules of FileLineReader (file_line_reader-inl.hpp:12-19), the app parser's CHECK (yelp.cpp:62)
  * split rule of random_split_by_feature_group (data-inl.hpp:231-272)
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    from cdae_b200 import build
    build.build()
    import cdae_b200
    return cdae_b200


def test_golden_pairs_match_reference_loader(lib):
    g = np.load(os.path.join(HERE, "golden", "pairs_small.npz"))
    d = lib.Dataset(os.path.join(HERE, "golden", "pairs_small.txt"))
    assert (d.num_users, d.num_items) == (int(g["U"]), int(g["I"]))
    rp, col = d.csr("all")
    np.testing.assert_array_equal(rp, g["row_ptr"])
    np.testing.assert_array_equal(col, g["col"])
    first = open(os.path.join(HERE, "golden", "pairs_small.txt")).readline().split()
    assert d.raw_id(0, 0) == first[0] and d.raw_id(1, 0) == first[1]      # ids in first-seen order
    assert d.num_instances == sum(1 for l in open(os.path.join(HERE, "golden", "pairs_small.txt")) if l.strip())


def test_live_against_reference_loader(lib, tmp_path, oracle_built):
    orc = oracle_built
    if not orc.have_reference():
        pytest.skip("needs oracle/_ref (the verbatim reference)")
    rng = np.random.default_rng(3)
    lines = ["%d %d" % (u, i) for u, i in zip(rng.integers(1000, 1400, 6000), rng.integers(0, 900, 6000))]
    p = tmp_path / "pairs.txt"
    p.write_text("\n".join(lines) + "\n")
    d = lib.Dataset(p)
    rp, col = d.csr("all")
    import ctypes as C
    L = orc._ref()
    h = L.ref_create((C.c_double * 4)(0.01, 0.1, 0.5, 0.0), (C.c_int32 * 11)(5, 4, 1, 1, 1, 0, 1, 0, 1, 0, 0), str(p).encode())
    try:
        assert (L.ref_num_users(h), L.ref_num_items(h)) == (d.num_users, d.num_items)
        buf = np.zeros(d.num_items, np.int64)
        for u in range(d.num_users):
            n = L.ref_user_items(h, u, buf.ctypes.data_as(orc.i64p), d.num_items)
            assert np.sort(buf[:n]).tolist() == col[rp[u]:rp[u + 1]].tolist()
    finally:
        L.ref_destroy(h)


def test_tokenizer_and_line_rules(lib, tmp_path):
    from cdae_b200 import CdaeError
    # any character of the delimiter string separates; empty tokens are dropped (the reference's known
    # answer: "12%&123124#$%&*,asdj#lwei#$" split on "#$" -> 3 tokens); empty lines are skipped and
    # NOT counted, so the header is the first non-empty line; CRLF tolerated
    p = tmp_path / "a.txt"
    p.write_text("\n\nuser#item\nalice#$#$x1\r\n\nbob$x2#\nalice##x2\n")
    d = lib.Dataset(p, delimiters="#$", skip_header=True)
    assert (d.num_users, d.num_items, d.num_instances) == (2, 2, 3)
    assert [d.raw_id(0, k) for k in range(2)] == ["alice", "bob"]
    assert [d.raw_id(1, k) for k in range(2)] == ["x1", "x2"]
    rp, col = d.csr()
    assert rp.tolist() == [0, 2, 3] and col.tolist() == [0, 1, 1]
    q = tmp_path / "b.txt"
    q.write_text("12%&123124#$%&*,asdj#lwei#$\n")            # three fields: the app's parser CHECK-aborts
    with pytest.raises(CdaeError):
        lib.Dataset(q, delimiters="#$")
    with pytest.raises(CdaeError):
        lib.Dataset(tmp_path / "missing.txt")


def test_split_rule(lib):
    d = lib.Dataset(os.path.join(HERE, "golden", "pairs_small.txt"))
    rp, col = d.csr("all")
    (trp, tcol), (erp, ecol) = d.random_split_by_feature_group(0.2, seed=5)
    lines = [l.split() for l in open(os.path.join(HERE, "golden", "pairs_small.txt")) if l.strip()]
    users = {}
    for u, _ in lines:
        users.setdefault(u, 0)
        users[u] += 1
    for u in range(d.num_users):
        n_inst = users[d.raw_id(0, u)]
        tr, te = set(tcol[trp[u]:trp[u + 1]].tolist()), set(ecol[erp[u]:erp[u + 1]].tolist())
        assert tr | te == set(col[rp[u]:rp[u + 1]].tolist())
        # floor(n * ratio) INSTANCES go to test (data-inl.hpp:252); duplicate pairs can shrink the set
        assert len(te) <= int(n_inst * 0.2) and len(tr) >= 1
    (trp2, tcol2), _ = d.random_split_by_feature_group(0.2, seed=5)
    np.testing.assert_array_equal(tcol, tcol2)
    (_, tcol3), _ = d.random_split_by_feature_group(0.2, seed=6)
    assert not np.array_equal(tcol, tcol3)
    n_test_total = int(sum(int(c * 0.2) for c in users.values()))
    assert 0 < len(ecol) <= n_test_total


def test_cache_round_trip(lib, tmp_path):
    """cdae_dataset_save / cdae_dataset_load (Data::save / Data::load, data.hpp:25-33, 52-60):
    ids, instances and the split survive a round trip; an unsplit data set stays unsplit."""
    from cdae_b200 import CdaeError
    src = os.path.join(HERE, "golden", "pairs_small.txt")
    d = lib.Dataset(src)
    p0 = tmp_path / "unsplit.cdaeds"
    d.save(p0)
    e = lib.Dataset.load(p0)
    assert (e.num_users, e.num_items, e.num_instances) == (d.num_users, d.num_items, d.num_instances)
    for a, b in zip(d.csr("all"), e.csr("all")):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(CdaeError):
        e.csr("train")                                           # no split was stored
    assert [e.raw_id(0, k) for k in range(e.num_users)] == [d.raw_id(0, k) for k in range(d.num_users)]
    assert [e.raw_id(1, k) for k in range(e.num_items)] == [d.raw_id(1, k) for k in range(d.num_items)]
    (trp, tcol), (erp, ecol) = d.random_split_by_feature_group(0.2, seed=5)
    p1 = tmp_path / "split.cdaeds"
    d.save(p1)
    e = lib.Dataset.load(p1)
    for which, want in (("all", d.csr("all")), ("train", (trp, tcol)), ("test", (erp, ecol))):
        got = e.csr(which)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
    # a fresh split of the loaded set with the same seed is the same split (the instances are in file order)
    (trp2, tcol2), _ = e.random_split_by_feature_group(0.2, seed=5)
    np.testing.assert_array_equal(trp2, trp)
    np.testing.assert_array_equal(tcol2, tcol)


def test_cache_rejects_damaged_files(lib, tmp_path):
    from cdae_b200 import CdaeError
    d = lib.Dataset(os.path.join(HERE, "golden", "pairs_small.txt"))
    d.random_split_by_feature_group(0.2, seed=5)
    p = tmp_path / "ok.cdaeds"
    d.save(p)
    blob = p.read_bytes()
    assert blob[:8] == b"CDAEDS01"
    cases_ = {
        "magic": b"XDAEDS01" + blob[8:],
        "version": blob[:8] + (99).to_bytes(4, "little") + blob[12:],
        "cut_ids": blob[:40],
        "cut_instances": blob[:len(blob) // 2],
        "cut_tail": blob[:-5],
        "huge_count": blob[:12] + (2**62).to_bytes(8, "little") + blob[20:],
        "empty": b"",
    }
    for name, b in cases_.items():
        q = tmp_path / (name + ".cdaeds")
        q.write_bytes(b)
        with pytest.raises(CdaeError):
            lib.Dataset.load(q)
    # a column id pushed out of range in the stored test part (the last 4 bytes of the file are its last entry)
    bad = bytearray(blob)
    bad[-4:] = (2**30).to_bytes(4, "little")
    q = tmp_path / "bad_col.cdaeds"
    q.write_bytes(bytes(bad))
    with pytest.raises(CdaeError):
        lib.Dataset.load(q)
    with pytest.raises(CdaeError):
        lib.Dataset.load(tmp_path / "missing.cdaeds")
