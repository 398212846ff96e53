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


def test_argument_checks_need_no_gpu(built):
    """cdae_create range-checks the configuration (q in [0,1], finite lambda / lr, ...) — before any CUDA
    call, so it is testable here; the group constructor checks its device count the same way."""
    import ctypes as C
    import numpy as np
    from cdae_b200 import _lib
    L = _lib.lib()
    rp = np.array([0, 2, 4], np.int64)
    col = np.array([0, 1, 1, 2], np.int32)
    h = C.c_void_p()

    def create(**kw):
        c = _lib.Config()
        assert L.cdae_config_default(C.byref(c)) == 0
        c.loss_type = _lib.LOSS["CE"]
        for k, v in kw.items():
            setattr(c, k, v)
        return L.cdae_create(C.byref(c), 2, 3, rp.ctypes.data_as(_lib.i64p), col.ctypes.data_as(_lib.i32p), C.byref(h))

    for bad in (dict(corruption_ratio=1.5), dict(corruption_ratio=-0.1), dict(corruption_ratio=float("nan")),
                dict(lambda_=float("inf")), dict(lambda_=-1.0), dict(learn_rate=float("nan")), dict(beta=-1.0),
                dict(num_dim=0), dict(num_dim=4096), dict(num_neg=-1), dict(loss_type=99)):
        assert create(**bad) == -1, bad                                    # CDAE_E_INVALID
        assert L.cdae_last_error()
    bad_col = np.array([0, 1, 1, 7], np.int32)                              # item id 7 outside [0, 3)
    c = _lib.Config()
    L.cdae_config_default(C.byref(c))
    assert L.cdae_create(C.byref(c), 2, 3, rp.ctypes.data_as(_lib.i64p), bad_col.ctypes.data_as(_lib.i32p), C.byref(h)) == -1
    g = C.c_void_p()
    for n in (0, 9):
        assert L.cdae_group_create(C.byref(c), 2, 3, rp.ctypes.data_as(_lib.i64p), col.ctypes.data_as(_lib.i32p), None, n, C.byref(g)) == -1
