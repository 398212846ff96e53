"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/cdae_b200.h
declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from cdae_b200 import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cdae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cdae_[a-z0-9_]+)\s*\(", src)))


def test_header_is_plain_c(tmp_path):
    """The boundary header must compile as C (no C++ / torch types in the signatures)."""
    c = tmp_path / "t.c"
    c.write_text('#include "cdae_b200.h"\nint main(void){cdae_config_t c; return cdae_config_default(&c) && 0;}\n')
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only",
                           "-I", os.path.join(ROOT, "include"), str(c)])


def test_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libcdae_b200.so does not export " + n


def test_bindings_cover_the_header(built):
    from cdae_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    L = _lib.lib()
    assert L.cdae_abi_version() == 1


def test_config_defaults_match_reference_struct(built):
    """cdae.hpp:14-30"""
    from cdae_b200 import CDAEConfig
    c = CDAEConfig()
    assert (c.lambda_, c.learn_rate, c.corruption_ratio, c.beta) == (0.01, 0.1, 0.5, 0.0)
    assert (c.num_dim, c.num_neg, c.num_corruptions) == (10, 5, 1)
    assert c.using_adagrad and c.user_factor and c.scaled
    assert not (c.asymmetric or c.linear or c.linear_function or c.tanh)
    assert c.loss == "LOGISTIC"


def test_sass_has_vector_reductions(built):
    """The scatter path must compile to 16-byte L2 reductions (REDG ... F32x4)."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    assert "REDG.E.ADD.F32x4" in sass
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout


def test_sass_is_blackwell_native(built):
    """The tensor-core paths must compile to tcgen05 / TMA / TMEM instructions (no mma.sync
    fallback): UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store (.MULTICAST: the
    2-CTA cluster variants), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit; the full-decode epilogue
    uses packed fp32x2 FMAs (FFMA2) and both SFU operations; the peer-memory all-reduce uses
    system-scope acquire / release accesses."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    for op in ("UTCHMMA", "UTMALDG.2D", "UTMALDG.2D.MULTICAST", "UTMASTG.2D", "LDTM.x32", "UTCBAR", "UTCBAR.MULTICAST",
               "FFMA2", "MUFU.EX2", "MUFU.RCP", "UCGABAR_ARV"):
        assert op in sass, op
    assert "HMMA" not in sass.replace("UTCHMMA", "")            # no warp-level mma.sync anywhere
    assert ".STRONG.SYS" in sass                                # ld.acquire.sys / st.release.sys of the p2p flags
