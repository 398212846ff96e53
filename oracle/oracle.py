"""ctypes bindings for the parity checker (TEST INFRASTRUCTURE, not product code).

* ``Oracle``    — oracle/_build/libcdae_oracle.so, the plain-C fp64 restatement of
                  /root/reference/src/model/recsys/cdae.hpp (built by ``make -C oracle``).
* ``Reference`` — oracle/_ref/libcdae_ref.so, the VERBATIM reference headers compiled
                  against the compat/ stand-ins (built by ``make -C oracle ref`` where
                  /root/reference exists; the prebuilt .so travels to the GPU box).
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libcdae_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcdae_ref.so")
REF_V3_SO = os.path.join(HERE, "_ref", "libcdae_ref_v3.so")   # same sources, -march=x86-64-v3 (bench baseline only)

LOSS = {"SQUARE": 0, "LOGISTIC": 1, "LOG": 2, "HINGE": 3, "SQUARED_HINGE": 4, "CE": 5,
        "CROSS_ENTROPY": 5, "LOGM": 6}
PARAMS = ["W", "V", "Wu", "b", "b_prime", "Uu",
          "W_ag", "V_ag", "Wu_ag", "b_ag", "b_prime_ag", "Uu_ag"]
PARAM_ID = {n: i for i, n in enumerate(PARAMS)}

i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


def build(ref=True):
    """Compile the checker (and, where /root/reference exists, the reference driver)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref and os.path.isdir("/root/reference/src/model/recsys"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "ref_v3"])


def default_config(**kw):
    """CDAEConfig defaults (cdae.hpp:13-31) with loss spelled as in apps/yelp (yelp.cpp:183-195)."""
    cfg = dict(lambda_=0.01, learn_rate=0.1, corruption_ratio=0.5, beta=0.0, loss="LOGISTIC",
               num_dim=10, num_neg=5, num_corruptions=1, using_adagrad=True, asymmetric=False,
               user_factor=True, linear=False, scaled=True, linear_function=False, tanh=False)
    for k, v in kw.items():
        if k not in cfg:
            raise KeyError(k)
        cfg[k] = v
    return cfg


class _OrcConfig(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("learn_rate", C.c_double),
                ("corruption_ratio", C.c_double), ("beta", C.c_double),
                ("loss_type", C.c_int32), ("num_dim", C.c_int32), ("num_neg", C.c_int32),
                ("num_corruptions", C.c_int32), ("using_adagrad", C.c_int32),
                ("asymmetric", C.c_int32), ("user_factor", C.c_int32), ("linear", C.c_int32),
                ("scaled", C.c_int32), ("linear_function", C.c_int32), ("tanh_act", C.c_int32)]


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a, t):
    return a.ctypes.data_as(t)


_orc_lib = None


def _orc():
    global _orc_lib
    if _orc_lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_OrcConfig), C.c_int64, C.c_int64, i64p, i32p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_param.restype = f64p
        L.orc_param.argtypes = [C.c_void_p, C.c_int, i64p, i64p]
        L.orc_init_params.argtypes = [C.c_void_p, C.c_uint64]
        for f in (L.orc_loss_gradient, L.orc_loss_evaluate):
            f.restype = C.c_double
            f.argtypes = [C.c_int, C.c_double, C.c_double]
        L.orc_hidden.argtypes = [C.c_void_p, C.c_int64, i64p, C.c_int64, C.c_double, f64p]
        L.orc_output.restype = C.c_double
        L.orc_output.argtypes = [C.c_void_p, f64p, C.c_int64]
        L.orc_step_sequential.argtypes = [C.c_void_p, C.c_int64, i64p, C.c_int64, i64p, C.c_int64,
                                          i64p]
        L.orc_step_frozen.argtypes = [C.c_void_p, C.c_int64, i64p, i64p, i64p, i64p, i64p, f64p]
        L.orc_shard_gradients.argtypes = [C.c_void_p, C.c_int64, i64p, i64p, i64p, i64p, i64p, f64p,
                                          f64p]
        L.orc_apply_dense.argtypes = [C.c_void_p, f64p, C.c_int]
        L.orc_step_frozen_full.argtypes = [C.c_void_p, C.c_int64, i64p, i64p, i64p, C.c_int, f64p]
        L.orc_shard_gradients_full.argtypes = [C.c_void_p, C.c_int64, i64p, i64p, i64p, C.c_int, f64p, f64p]
        L.orc_train_epoch_full.restype = C.c_double
        L.orc_train_epoch_full.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int64,
                                           C.c_int64, C.c_int]
        L.orc_recommend.restype = C.c_int
        L.orc_recommend.argtypes = [C.c_void_p, C.c_int64, C.c_int64, i64p, f64p]
        L.orc_data_loss.restype = C.c_double
        L.orc_data_loss.argtypes = [C.c_void_p, u8p]
        L.orc_penalty_loss.restype = C.c_double
        L.orc_penalty_loss.argtypes = [C.c_void_p]
        L.orc_user_representations.argtypes = [C.c_void_p, f64p]
        L.orc_evaluate_rec_list.argtypes = [i64p, C.c_int64, i64p, C.c_int64, f64p]
        L.orc_topn_evaluate.restype = C.c_int64
        L.orc_topn_evaluate.argtypes = [C.c_void_p, i64p, i32p, f64p]
        L.orc_philox4x32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.POINTER(C.c_uint32)]
        L.orc_sample_keep.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, u8p]
        L.orc_sample_negatives.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, i64p]
        L.orc_train_epoch.restype = C.c_double
        L.orc_train_epoch.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_int64]
        L.orc_train_epoch_hogwild.restype = C.c_double
        L.orc_train_epoch_hogwild.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64,
                                              C.c_int64, C.c_int]
        _orc_lib = L
    return _orc_lib


def philox4x32(seed, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    _orc().orc_philox4x32(seed, c0, c1, c2, c3, out)
    return [int(x) for x in out]


def loss_gradient(loss, pred, truth):
    return _orc().orc_loss_gradient(LOSS[loss] if isinstance(loss, str) else loss, pred, truth)


def loss_evaluate(loss, pred, truth):
    return _orc().orc_loss_evaluate(LOSS[loss] if isinstance(loss, str) else loss, pred, truth)


def evaluate_rec_list(lst, test_items):
    lst = _as(lst, np.int64)
    t = _as(test_items, np.int64)
    out = np.zeros(8)
    _orc().orc_evaluate_rec_list(_p(lst, i64p), len(lst), _p(t, i64p), len(t), _p(out, f64p))
    return out


class Oracle:
    """The plain-C restatement."""

    def __init__(self, cfg, U, I, row_ptr, col):
        self.cfg = dict(cfg)
        self.U, self.I, self.K = int(U), int(I), int(cfg["num_dim"])
        self.row_ptr = _as(row_ptr, np.int64)
        self.col = _as(col, np.int32)
        c = _OrcConfig(cfg["lambda_"], cfg["learn_rate"], cfg["corruption_ratio"], cfg["beta"],
                       LOSS[cfg["loss"]], cfg["num_dim"], cfg["num_neg"], cfg["num_corruptions"],
                       int(cfg["using_adagrad"]), int(cfg["asymmetric"]), int(cfg["user_factor"]),
                       int(cfg["linear"]), int(cfg["scaled"]), int(cfg["linear_function"]),
                       int(cfg["tanh"]))
        self._L = _orc()
        self._h = self._L.orc_create(C.byref(c), self.U, self.I, _p(self.row_ptr, i64p),
                                     _p(self.col, i32p))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(self._h)
            self._h = None

    def param(self, name):
        """Writable numpy VIEW of a parameter block."""
        r, c = C.c_int64(), C.c_int64()
        p = self._L.orc_param(self._h, PARAM_ID[name], C.byref(r), C.byref(c))
        n = r.value * c.value
        if n == 0:
            return np.zeros((r.value, c.value))
        a = np.ctypeslib.as_array(p, shape=(n,))
        return a.reshape(r.value, c.value) if c.value > 1 else a

    def set_params(self, params):
        for k, v in params.items():
            self.param(k)[...] = np.asarray(v, dtype=np.float64).reshape(self.param(k).shape)

    def get_params(self):
        return {k: self.param(k).copy() for k in PARAMS}

    def init_params(self, seed):
        self._L.orc_init_params(self._h, seed)

    def hidden(self, uid, items, scale=1.0):
        items = _as(items, np.int64)
        z = np.zeros(self.K)
        self._L.orc_hidden(self._h, uid, _p(items, i64p), len(items), scale, _p(z, f64p))
        return z

    def output(self, z, item):
        z = _as(z, np.float64)
        return self._L.orc_output(self._h, _p(z, f64p), item)

    def step_sequential(self, uid, in_items, negs, out_order=None):
        a = _as(in_items, np.int64)
        n = _as(negs, np.int64)
        o = None if out_order is None else _as(out_order, np.int64)
        self._L.orc_step_sequential(self._h, uid, _p(a, i64p), len(a), _p(n, i64p), len(n),
                                    None if o is None else _p(o, i64p))

    def step_frozen(self, uids, in_lists, neg_lists):
        uids = _as(uids, np.int64)
        in_ptr = np.zeros(len(uids) + 1, np.int64)
        neg_ptr = np.zeros(len(uids) + 1, np.int64)
        in_ptr[1:] = np.cumsum([len(x) for x in in_lists])
        neg_ptr[1:] = np.cumsum([len(x) for x in neg_lists])
        ins = _as(np.concatenate([np.asarray(x, np.int64) for x in in_lists] + [np.zeros(0, np.int64)]), np.int64)
        ngs = _as(np.concatenate([np.asarray(x, np.int64) for x in neg_lists] + [np.zeros(0, np.int64)]), np.int64)
        ls = C.c_double(0)
        self._L.orc_step_frozen(self._h, len(uids), _p(uids, i64p), _p(in_ptr, i64p),
                                _p(ins, i64p), _p(neg_ptr, i64p), _p(ngs, i64p), C.byref(ls))
        return ls.value

    def step_frozen_full(self, uids, in_lists, rounding=0):
        """H12: frozen minibatch with every user's output set = all items (no reference
        function; == step_frozen with negatives = every non-positive once).  rounding=1 restates
        the bf16 operand rounding of the tensor-core path."""
        uids = _as(uids, np.int64)
        in_ptr = np.zeros(len(uids) + 1, np.int64)
        in_ptr[1:] = np.cumsum([len(x) for x in in_lists])
        ins = _as(np.concatenate([np.asarray(x, np.int64) for x in in_lists] + [np.zeros(0, np.int64)]), np.int64)
        ls = C.c_double(0)
        self._L.orc_step_frozen_full(self._h, len(uids), _p(uids, i64p), _p(in_ptr, i64p), _p(ins, i64p),
                                     int(rounding), C.byref(ls))
        return ls.value

    def shard_gradients_full(self, uids, in_lists, dense, rounding=0):
        uids = _as(uids, np.int64)
        in_ptr = np.zeros(len(uids) + 1, np.int64)
        in_ptr[1:] = np.cumsum([len(x) for x in in_lists])
        ins = _as(np.concatenate([np.asarray(x, np.int64) for x in in_lists] + [np.zeros(0, np.int64)]), np.int64)
        assert dense.dtype == np.float64 and dense.flags.c_contiguous and dense.size == self.dense_grad_size()
        ls = C.c_double(0)
        self._L.orc_shard_gradients_full(self._h, len(uids), _p(uids, i64p), _p(in_ptr, i64p), _p(ins, i64p),
                                         int(rounding), C.byref(ls), _p(dense, f64p))
        return ls.value

    def train_epoch_full(self, seed, epoch, batch_users, rounding=0, u0=0, u1=None):
        u1 = self.U if u1 is None else u1
        return self._L.orc_train_epoch_full(self._h, seed, epoch, batch_users, u0, u1, int(rounding))

    def dense_grad_size(self):
        """[gW | gV (asymmetric) | gb' | gb] — the buffer the GPU path all-reduces."""
        K = self.cfg["num_dim"]
        return self.I * K * (2 if self.cfg["asymmetric"] else 1) + self.I + K

    def shard_gradients(self, uids, in_lists, neg_lists, dense):
        """Frozen-parameter gradients of a minibatch SHARD: item side added to `dense`,
        user-private rows updated in place.  Returns the shard's loss sum."""
        uids = _as(uids, np.int64)
        in_ptr = np.zeros(len(uids) + 1, np.int64)
        neg_ptr = np.zeros(len(uids) + 1, np.int64)
        in_ptr[1:] = np.cumsum([len(x) for x in in_lists])
        neg_ptr[1:] = np.cumsum([len(x) for x in neg_lists])
        ins = _as(np.concatenate([np.asarray(x, np.int64) for x in in_lists] + [np.zeros(0, np.int64)]), np.int64)
        ngs = _as(np.concatenate([np.asarray(x, np.int64) for x in neg_lists] + [np.zeros(0, np.int64)]), np.int64)
        assert dense.dtype == np.float64 and dense.flags.c_contiguous and dense.size == self.dense_grad_size()
        ls = C.c_double(0)
        self._L.orc_shard_gradients(self._h, len(uids), _p(uids, i64p), _p(in_ptr, i64p),
                                    _p(ins, i64p), _p(neg_ptr, i64p), _p(ngs, i64p), C.byref(ls),
                                    _p(dense, f64p))
        return ls.value

    def apply_dense(self, dense, any_steps=True):
        self._L.orc_apply_dense(self._h, _p(dense, f64p), int(any_steps))

    def recommend(self, uid, topk=10):
        ids = np.zeros(topk, np.int64)
        sc = np.zeros(topk)
        rc = self._L.orc_recommend(self._h, uid, topk, _p(ids, i64p), _p(sc, f64p))
        if rc != 0:
            raise RuntimeError("fewer than topk unrated items (reference CHECK_EQ, cdae.hpp:187)")
        return ids, sc

    def data_loss(self, keep=None):
        if keep is None:
            return self._L.orc_data_loss(self._h, None)
        k = _as(keep, np.uint8)
        return self._L.orc_data_loss(self._h, _p(k, u8p))

    def penalty_loss(self):
        return self._L.orc_penalty_loss(self._h)

    def user_representations(self):
        out = np.zeros((self.U, self.K))
        self._L.orc_user_representations(self._h, _p(out, f64p))
        return out

    def topn_evaluate(self, test_row_ptr, test_col):
        rp = _as(test_row_ptr, np.int64)
        cl = _as(test_col, np.int32)
        out = np.zeros(8)
        n = self._L.orc_topn_evaluate(self._h, _p(rp, i64p), _p(cl, i32p), _p(out, f64p))
        return out, n

    def sample_keep(self, seed, pass_, uid):
        n = int(self.row_ptr[uid + 1] - self.row_ptr[uid])
        k = np.zeros(max(n, 1), np.uint8)
        self._L.orc_sample_keep(self._h, seed, pass_, uid, _p(k, u8p))
        return k[:n]

    def sample_negatives(self, seed, pass_, uid):
        n = int(self.row_ptr[uid + 1] - self.row_ptr[uid]) * self.cfg["num_neg"]
        g = np.zeros(max(n, 1), np.int64)
        self._L.orc_sample_negatives(self._h, seed, pass_, uid, _p(g, i64p))
        return g[:n]

    def train_epoch(self, seed, epoch, batch_users=1, u0=0, u1=None):
        u1 = self.U if u1 is None else u1
        return self._L.orc_train_epoch(self._h, seed, epoch, batch_users, u0, u1)

    def train_epoch_hogwild(self, seed, epoch, n_threads, u0=0, u1=None):
        u1 = self.U if u1 is None else u1
        return self._L.orc_train_epoch_hogwild(self._h, seed, epoch, u0, u1, n_threads)


_ref_libs = {}


def have_reference():
    return os.path.exists(REF_SO)


def _ref(so=None):
    """The reference driver library (default build, or another build of the same sources)."""
    so = so or REF_SO
    if so not in _ref_libs:
        _ref_libs[so] = _ref_load(so)
    return _ref_libs[so]


_ref_libs = {}


def _ref_load(path):
    _ref_lib = None
    if _ref_lib is None:
        L = C.CDLL(path)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [f64p, i32p, C.c_char_p]
        L.ref_destroy.argtypes = [C.c_void_p]
        for f in (L.ref_num_users, L.ref_num_items, L.ref_pending_negatives):
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p]
        L.ref_seed.argtypes = [C.c_uint64, C.c_uint32]
        L.ref_set_num_thread.argtypes = [C.c_int32]
        L.ref_param_shape.argtypes = [C.c_void_p, C.c_int, i64p, i64p]
        L.ref_set_param.argtypes = [C.c_void_p, C.c_int, f64p, C.c_int64]
        L.ref_get_param.argtypes = [C.c_void_p, C.c_int, f64p, C.c_int64]
        L.ref_user_items.restype = C.c_int64
        L.ref_user_items.argtypes = [C.c_void_p, C.c_int64, i64p, C.c_int64]
        L.ref_push_negatives.argtypes = [C.c_void_p, i64p, C.c_int64]
        L.ref_train_one_user.argtypes = [C.c_void_p, C.c_int64, i64p, C.c_int64]
        L.ref_train_one_iteration.restype = C.c_double
        L.ref_train_one_iteration.argtypes = [C.c_void_p]
        L.ref_train_user_range.restype = C.c_double
        L.ref_train_user_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.ref_corrupt.restype = C.c_int64
        L.ref_corrupt.argtypes = [C.c_void_p, C.c_int64, C.c_double, i64p, C.c_int64]
        L.ref_hidden.argtypes = [C.c_void_p, C.c_int64, i64p, C.c_int64, C.c_double, f64p]
        L.ref_output.restype = C.c_double
        L.ref_output.argtypes = [C.c_void_p, f64p, C.c_int64]
        L.ref_recommend.argtypes = [C.c_void_p, C.c_int64, C.c_int64, i64p]
        L.ref_user_representations.argtypes = [C.c_void_p, f64p]
        L.ref_data_loss.restype = C.c_double
        L.ref_data_loss.argtypes = [C.c_void_p]
        L.ref_penalty_loss.restype = C.c_double
        L.ref_penalty_loss.argtypes = [C.c_void_p]
        for f in (L.ref_loss_gradient, L.ref_loss_evaluate):
            f.restype = C.c_double
            f.argtypes = [C.c_int32, C.c_double, C.c_double]
        L.ref_evaluate_rec_list.argtypes = [i64p, C.c_int64, i64p, C.c_int64, f64p]
        L.ref_topn_evaluate.restype = C.c_int
        L.ref_topn_evaluate.argtypes = [C.c_void_p, C.c_char_p, f64p]
        _ref_lib = L
    return _ref_lib


def ref_loss_gradient(loss, pred, truth):
    return _ref().ref_loss_gradient(LOSS[loss], pred, truth)


def ref_loss_evaluate(loss, pred, truth):
    return _ref().ref_loss_evaluate(LOSS[loss], pred, truth)


def ref_evaluate_rec_list(lst, test_items):
    lst = _as(lst, np.int64)
    t = _as(test_items, np.int64)
    out = np.zeros(8)
    _ref().ref_evaluate_rec_list(_p(lst, i64p), len(lst), _p(t, i64p), len(t), _p(out, f64p))
    return out


def write_pairs(path, row_ptr, col):
    """CSR -> the "user item" text the reference's RECSYS loader reads."""
    row_ptr = np.asarray(row_ptr)
    users = np.repeat(np.arange(len(row_ptr) - 1), np.diff(row_ptr))
    np.savetxt(path, np.stack([users, np.asarray(col)], 1), fmt="%d")


class Reference:
    """The verbatim reference CDAE (single-threaded fp64), driven through ref_driver.cpp.

    The reference indexes users / items in FIRST-SEEN order of the text file
    (instance-inl.hpp:22-37).  ``identity_ids`` asserts that the CSR handed in is already
    numbered that way (users ascending, items numbered by first appearance), so ids agree
    with the oracle's; use ``first_seen_relabel`` to build such a CSR.
    """

    def __init__(self, cfg, U, I, row_ptr, col, quiet=True, so=None):
        if quiet:
            os.environ.setdefault("GLOG_minloglevel", "1")
        self.cfg = dict(cfg)
        self.U, self.I, self.K = int(U), int(I), int(cfg["num_dim"])
        self._L = _ref(so)
        d = _as([cfg["lambda_"], cfg["learn_rate"], cfg["corruption_ratio"], cfg["beta"]], np.float64)
        i = _as([LOSS[cfg["loss"]], cfg["num_dim"], cfg["num_neg"], cfg["num_corruptions"],
                 cfg["using_adagrad"], cfg["asymmetric"], cfg["user_factor"], cfg["linear"],
                 cfg["scaled"], cfg["linear_function"], cfg["tanh"]], np.int32)
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            path = f.name
        try:
            write_pairs(path, row_ptr, col)
            self._h = self._L.ref_create(_p(d, f64p), _p(i, i32p), path.encode())
        finally:
            os.unlink(path)
        assert self._L.ref_num_users(self._h) == self.U, "users must all appear, in ascending order"
        assert self._L.ref_num_items(self._h) == self.I, "every item id must appear in the data"

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_destroy(self._h)
            self._h = None

    @staticmethod
    def seed(mt_seed, c_seed, so=None):
        _ref(so).ref_seed(mt_seed, c_seed)

    def shape(self, name):
        r, c = C.c_int64(), C.c_int64()
        self._L.ref_param_shape(self._h, PARAM_ID[name], C.byref(r), C.byref(c))
        return r.value, c.value

    def set_params(self, params):
        for k, v in params.items():
            r, c = self.shape(k)
            if r * c == 0:
                continue
            a = _as(np.asarray(v, np.float64).reshape(-1), np.float64)
            rc = self._L.ref_set_param(self._h, PARAM_ID[k], _p(a, f64p), a.size)
            assert rc == 0, (k, rc, a.size, r, c)

    def get_param(self, name):
        r, c = self.shape(name)
        a = np.zeros(r * c)
        if a.size:
            assert self._L.ref_get_param(self._h, PARAM_ID[name], _p(a, f64p), a.size) == 0
        return a.reshape(r, c) if c > 1 else a

    def get_params(self):
        return {k: self.get_param(k) for k in PARAMS}

    def user_items_order(self, uid):
        n = self._L.ref_user_items(self._h, uid, None, 0)
        out = np.zeros(max(n, 1), np.int64)
        self._L.ref_user_items(self._h, uid, _p(out, i64p), n)
        return out[:n]

    def train_one_user(self, uid, in_items, negs):
        a = _as(in_items, np.int64)
        n = _as(negs, np.int64)
        self._L.ref_push_negatives(self._h, _p(n, i64p), len(n))
        self._L.ref_train_one_user(self._h, uid, _p(a, i64p), len(a))
        assert self._L.ref_pending_negatives(self._h) == 0, "negative count != n_u * num_neg"

    def train_one_iteration(self):
        return self._L.ref_train_one_iteration(self._h)

    def train_user_range(self, u0, u1):
        return self._L.ref_train_user_range(self._h, u0, u1)

    def corrupt(self, uid, ratio):
        out = np.zeros(self.I, np.int64)
        n = self._L.ref_corrupt(self._h, uid, ratio, _p(out, i64p), len(out))
        return out[:n]

    def hidden(self, uid, items, scale=1.0):
        items = _as(items, np.int64)
        z = np.zeros(self.K)
        self._L.ref_hidden(self._h, uid, _p(items, i64p), len(items), scale, _p(z, f64p))
        return z

    def output(self, z, item):
        z = _as(z, np.float64)
        return self._L.ref_output(self._h, _p(z, f64p), item)

    def recommend(self, uid, topk=10):
        out = np.zeros(topk, np.int64)
        self._L.ref_recommend(self._h, uid, topk, _p(out, i64p))
        return out

    def user_representations(self):
        out = np.zeros((self.U, self.K))
        self._L.ref_user_representations(self._h, _p(out, f64p))
        return out

    def data_loss(self):
        return self._L.ref_data_loss(self._h)

    def penalty_loss(self):
        return self._L.ref_penalty_loss(self._h)

    def topn_evaluate(self, test_row_ptr, test_col):
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            path = f.name
        try:
            write_pairs(path, test_row_ptr, test_col)
            out = np.zeros(8)
            assert self._L.ref_topn_evaluate(self._h, path.encode(), _p(out, f64p)) == 0
        finally:
            os.unlink(path)
        return out


def first_seen_relabel(row_ptr, col):
    """Renumber items by first appearance in (user-major) file order, so that the reference's
    string->index maps (instance-inl.hpp:22-37) are the identity.  Rows are re-sorted."""
    col = np.asarray(col, np.int64)
    _, first = np.unique(col, return_index=True)
    order = col[np.sort(first)]          # item ids in first-seen order
    new_id = np.empty(col.max() + 1, np.int64)
    new_id[:] = -1
    new_id[order] = np.arange(len(order))
    return new_id[col].astype(np.int32), len(order)
