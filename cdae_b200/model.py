"""Python mirror of `class CDAE` (reference: src/model/recsys/cdae.hpp) on top of the C ABI.

Same method names, argument meaning and error behaviour as the reference class so that parity
tests read like tests of the reference:

  reference (cdae.hpp)                                  here
  ----------------------------------------------------  -----------------------------------------
  CDAEConfig :13-31                                     CDAEConfig (same fields / defaults)
  CDAE::reset(const Data&) :109-134                     CDAE.reset(U, I, row_ptr, col)
  CDAE::train_one_iteration(const Data&) :136-146       CDAE.train_one_iteration()
  CDAE::train_one_user_corruption(uid,in,out) :198      CDAE.train_one_user_corruption / train_users
  CDAE::get_hidden_values(uid, set, scale) :373-416     CDAE.get_hidden_values
  CDAE::get_user_representations() :148-159             CDAE.get_user_representations
  CDAE::data_loss / penalty_loss / current_loss :78-107 same names
  pre_recommend() hook + CDAE::recommend :162-196       CDAE.pre_recommend, CDAE.recommend

The reference aborts through glog CHECK on misuse; here the same conditions raise CdaeError.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CdaeError, EpochStats, PARAM_ID, PARAMS  # noqa: F401


class CDAEConfig:
    """libcf::CDAEConfig (cdae.hpp:13-31) + the device options of struct cdae_config."""

    def __init__(self, **kw):
        c = _lib.Config()
        _lib.check(_lib.lib().cdae_config_default(C.byref(c)))
        self.lambda_ = c.lambda_
        self.learn_rate = c.learn_rate
        self.corruption_ratio = c.corruption_ratio
        self.beta = c.beta
        self.loss = "LOGISTIC"          # struct default (aborts for CDAE): set CE / SQUARE
        self.num_dim = c.num_dim
        self.num_neg = c.num_neg
        self.num_corruptions = c.num_corruptions
        self.using_adagrad = bool(c.using_adagrad)
        self.asymmetric = bool(c.asymmetric)
        self.user_factor = bool(c.user_factor)
        self.linear = bool(c.linear)
        self.scaled = bool(c.scaled)
        self.linear_function = bool(c.linear_function)
        self.tanh = bool(c.tanh_act)
        self.batch_users = 0
        self.device = 0
        self.full_decode = False        # H12: decode against all items on tcgen05 (no reference function)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)

    def to_c(self):
        c = _lib.Config()
        _lib.check(_lib.lib().cdae_config_default(C.byref(c)))
        c.lambda_, c.learn_rate = self.lambda_, self.learn_rate
        c.corruption_ratio, c.beta = self.corruption_ratio, self.beta
        c.loss_type = _lib.LOSS[self.loss] if isinstance(self.loss, str) else int(self.loss)
        c.num_dim, c.num_neg, c.num_corruptions = self.num_dim, self.num_neg, self.num_corruptions
        c.using_adagrad, c.asymmetric = int(self.using_adagrad), int(self.asymmetric)
        c.user_factor, c.linear, c.scaled = int(self.user_factor), int(self.linear), int(self.scaled)
        c.linear_function, c.tanh_act = int(self.linear_function), int(self.tanh)
        c.batch_users, c.device = int(self.batch_users), int(self.device)
        c.full_decode = int(self.full_decode)
        return c


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, t):
    return a.ctypes.data_as(t)


class CDAE:
    def __init__(self, config=None):
        self.config = config or CDAEConfig()
        self._h = C.c_void_p()
        self._L = _lib.lib()
        self.U = self.I = 0
        self.last_stats = None
        self._topk = 0

    # -- lifetime ---------------------------------------------------------------------------
    def reset(self, U, I, row_ptr, col):
        """CDAE::reset: takes the user->items structure (as CSR, rows ascending) and allocates
        the parameters.  Weights are zero until init_params() / set_params()."""
        self.close()
        self.U, self.I = int(U), int(I)
        self.row_ptr = _arr(row_ptr, np.int64)
        self.col = _arr(col, np.int32)
        assert len(self.row_ptr) == self.U + 1
        c = self.config.to_c()
        _lib.check(self._L.cdae_create(C.byref(c), self.U, self.I, _ptr(self.row_ptr, _lib.i64p),
                                       _ptr(self.col, _lib.i32p), C.byref(self._h)))
        return self

    def close(self):
        for p in getattr(self, "_pinned", []):
            self._L.cdae_host_free(p)
        self._pinned = []
        if self._h:
            self._L.cdae_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_params(self, seed):
        self._topk = 0
        _lib.check(self._L.cdae_init_params(self._h, seed))

    # -- parameters -------------------------------------------------------------------------
    def param_shape(self, name):
        r, c = C.c_int64(), C.c_int64()
        _lib.check(self._L.cdae_param_shape(self._h, PARAM_ID[name], C.byref(r), C.byref(c)))
        return r.value, c.value

    def set_params(self, params):
        self._topk = 0
        for k, v in params.items():
            r, c = self.param_shape(k)
            if r * c == 0:
                continue
            a = _arr(np.asarray(v, np.float64).reshape(-1), np.float64)
            _lib.check(self._L.cdae_set_param(self._h, PARAM_ID[k], _ptr(a, _lib.f64p), a.size))

    def get_param(self, name):
        r, c = self.param_shape(name)
        a = np.zeros(r * c)
        if a.size:
            _lib.check(self._L.cdae_get_param(self._h, PARAM_ID[name], _ptr(a, _lib.f64p), a.size))
        return a.reshape(r, c) if c > 1 else a

    def _get_rows(self, name, rows):
        """Selected rows of a block (cdae_get_param_rows; collective for user blocks in a process group)."""
        r = _arr(rows, np.int64)
        _, c = self.param_shape(name)
        out = np.zeros((len(r), max(c, 1)))
        _lib.check(self._L.cdae_get_param_rows(self._h, PARAM_ID[name], _ptr(r, _lib.i64p), len(r), _ptr(out, _lib.f64p)))
        return out

    def get_params(self):
        return {k: self.get_param(k) for k in PARAMS}

    # -- training ---------------------------------------------------------------------------
    def train_one_iteration(self, seed=0, epoch=0, csr=None):
        """CDAE::train_one_iteration.  csr=(row_ptr, col) passes the training data from host
        memory on this call (the reference receives `const Data&` every epoch)."""
        st = EpochStats()
        if csr is None:
            _lib.check(self._L.cdae_train_epoch(self._h, seed, epoch, C.byref(st)))
        else:
            rp, cl = _arr(csr[0], np.int64), _arr(csr[1], np.int32)   # no copy when already int64 / int32
            if len(rp) != self.U + 1 or len(cl) < rp[-1]:
                raise CdaeError(-1, "csr must be (row_ptr[U+1] int64, col[nnz] int32)")
            _lib.check(self._L.cdae_train_epoch_csr(self._h, _ptr(rp, _lib.i64p), _ptr(cl, _lib.i32p),
                                                    seed, epoch, C.byref(st)))
        self.last_stats = st
        self._topk = 0                  # stored lists are stale
        return st

    def train_users(self, uids, keep_mask, negatives):
        """One frozen minibatch of distinct users with explicit masks / negatives."""
        u = _arr(uids, np.int64)
        k = _arr(keep_mask, np.uint8)
        n = None if negatives is None else _arr(negatives, np.int32)
        st = EpochStats()
        _lib.check(self._L.cdae_train_users(self._h, _ptr(u, _lib.i64p), len(u), _ptr(k, _lib.u8p),
                                            None if n is None else _ptr(n, _lib.i32p), C.byref(st)))
        self.last_stats = st
        self._topk = 0
        return st

    def train_one_user_corruption(self, uid, input_items, negatives):
        """CDAE::train_one_user_corruption(uid, input_set, output_set = the user's train row);
        input_items is the corrupted input set, negatives the n_u*num_neg sampled negatives."""
        row = self.col[self.row_ptr[uid]:self.row_ptr[uid + 1]]
        keep = np.isin(row, np.asarray(input_items)).astype(np.uint8)
        if keep.sum() != len(set(np.asarray(input_items).tolist())):
            raise CdaeError(-1, "input_set must be a subset of the user's train items")
        return self.train_users([uid], keep, negatives)

    # -- forward ----------------------------------------------------------------------------
    def encode(self, uids, keep_mask=None, scale=1.0):
        u = _arr(uids, np.int64)
        z = np.zeros((len(u), self.config.num_dim), np.float32)
        k = None if keep_mask is None else _arr(keep_mask, np.uint8)
        _lib.check(self._L.cdae_encode(self._h, _ptr(u, _lib.i64p), len(u),
                                       None if k is None else _ptr(k, _lib.u8p), scale,
                                       _ptr(z, _lib.f32p)))
        return z

    def get_hidden_values(self, uid, item_set, scale=1.0):
        """CDAE::get_hidden_values(uid, item_set, scale); item_set must be a subset of the row."""
        row = self.col[self.row_ptr[uid]:self.row_ptr[uid + 1]]
        keep = np.isin(row, np.asarray(item_set, np.int64)).astype(np.uint8)
        return self.encode([uid], keep, scale)[0]

    def get_user_representations(self):
        out = np.zeros((self.U, self.config.num_dim), np.float32)
        step = 1 << 16
        for a in range(0, self.U, step):
            b = min(self.U, a + step)
            out[a:b] = self.encode(np.arange(a, b))
        return out

    # -- losses -----------------------------------------------------------------------------
    def data_loss(self, seed=0):
        v = C.c_double()
        _lib.check(self._L.cdae_data_loss(self._h, seed, C.byref(v)))
        return v.value

    def penalty_loss(self):
        v = C.c_double()
        _lib.check(self._L.cdae_penalty_loss(self._h, C.byref(v)))
        return v.value

    def current_loss(self, seed=0):
        """ModelBase::current_loss = data_loss + penalty_loss (model_base.hpp:29-32)."""
        return self.data_loss(seed) + self.penalty_loss()

    # -- recommendation ---------------------------------------------------------------------
    def pre_recommend(self, topk=10):
        """The reference's pre_recommend() hook (recsys_model_base.hpp:72): scores all users
        against all items on the device and stores every user's top-k list."""
        _lib.check(self._L.cdae_topn_build(self._h, topk))
        self._topk = topk

    def recommend(self, uid, topk=10):
        """CDAE::recommend(uid, topk, rated = train items of uid)."""
        if self._topk != topk:
            self.pre_recommend(topk)
        ids = np.zeros(topk, np.int64)
        sc = np.zeros(topk, np.float32)
        _lib.check(self._L.cdae_topn_lookup(self._h, uid, _ptr(ids, _lib.i64p), _ptr(sc, _lib.f32p)))
        return ids, sc

    def recommend_all(self, topk=10):
        if self._topk != topk:
            self.pre_recommend(topk)
        ids = np.zeros((self.U, topk), np.int64)
        sc = np.zeros((self.U, topk), np.float32)
        _lib.check(self._L.cdae_topn_fetch(self._h, _ptr(ids, _lib.i64p), _ptr(sc, _lib.f32p)))
        return ids, sc

    def topn_stats(self):
        """(path, verified_users, redone_users) of the last pre_recommend: path 1 = tcgen05
        candidates proven exact by the error bound, 0 = fp32 CUDA-core kernel."""
        p, a, b = C.c_int32(), C.c_int64(), C.c_int64()
        _lib.check(self._L.cdae_topn_stats(self._h, C.byref(p), C.byref(a), C.byref(b)))
        return p.value, a.value, b.value

    def topn_probe_items(self):
        """Items in the probe table of the last pre_recommend (0 = the first sweep started from -inf)."""
        n = C.c_int32()
        _lib.check(self._L.cdae_topn_probe_items(self._h, C.byref(n)))
        return n.value

    def topn_evaluate(self, test_row_ptr, test_col):
        """TOPN_Evaluation::evaluate on the stored lists: [P@1,P@5,P@10,R@1,R@5,R@10,MAP@5,MAP@10]."""
        rp = _arr(test_row_ptr, np.int64)
        cl = _arr(test_col, np.int32)
        out = np.zeros(8)
        n = C.c_int64()
        _lib.check(self._L.cdae_topn_evaluate(self._h, _ptr(rp, _lib.i64p), _ptr(cl, _lib.i32p),
                                              _ptr(out, _lib.f64p), C.byref(n)))
        return out, n.value

    # -- process group ----------------------------------------------------------------------
    @staticmethod
    def dist_unique_id():
        buf = (C.c_char * 128)()
        _lib.check(_lib.lib().cdae_dist_unique_id(buf))
        return bytes(buf)

    def dist_init(self, rank, world, unique_id):
        buf = C.create_string_buffer(unique_id, 128)
        _lib.check(self._L.cdae_dist_init(self._h, rank, world, buf))

    def dist_p2p_init(self, all_gather):
        """Switch the gradient all-reduce to the NVLink peer-memory kernel.  `all_gather(bytes) ->
        list of bytes in rank order` is the caller's collective (e.g. torch.distributed.all_gather_object)."""
        buf = (C.c_char * 256)()
        _lib.check(self._L.cdae_dist_p2p_export(self._h, buf))
        table = b"".join(all_gather(bytes(buf)))
        tb = C.create_string_buffer(table, len(table))
        _lib.check(self._L.cdae_dist_p2p_open(self._h, tb))

    def dist_mc_init(self, rank, world, all_gather):
        """Switch the combine step to the NVLS kernel (NVSwitch multicast, cdae_dist_mc_*).  `all_gather(obj)
        -> list in rank order` is the caller's collective.  Returns False — with nothing changed — where
        multicast is unavailable, so the caller can fall back to dist_p2p_init / NCCL; all ranks agree."""
        import os
        import socket
        fd, ok, srv, path = -1, True, None, None
        if rank == 0:
            v = C.c_int32(-1)
            rc = self._L.cdae_dist_mc_create(self._h, C.byref(v))
            ok, fd = rc == 0, v.value
            if ok:
                path = "/tmp/cdae_mc_%d.sock" % os.getpid()
                if os.path.exists(path):
                    os.unlink(path)
                srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                srv.bind(path)
                srv.listen(world)
        ok, path = all_gather((ok, path))[0]
        if not ok:
            return False
        if rank == 0:
            for _ in range(world - 1):                       # the multicast object travels as a file descriptor
                conn, _a = srv.accept()
                socket.send_fds(conn, [b"mc"], [fd])
                conn.close()
            srv.close()
            os.unlink(path)
        else:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(path)
            _msg, fds, _f, _a = socket.recv_fds(c, 16, 1)
            c.close()
            fd = fds[0]
        rc = self._L.cdae_dist_mc_attach(self._h, fd)
        if not all(all_gather(rc == 0)):
            # nothing on the training path has changed yet: every rank falls back together
            if rank != 0:
                os.close(fd)
            return False
        rc = self._L.cdae_dist_mc_bind(self._h)              # (the all_gather above was the barrier "everyone attached")
        if not all(all_gather(rc == 0)):
            raise CdaeError(rc, "cdae_dist_mc_bind failed on some rank: " + self._L.cdae_last_error().decode("utf-8", "replace"))
        if rank != 0:
            os.close(fd)
        return True

    def save(self, path):
        """Versioned binary checkpoint of every parameter block incl. AdaGrad state (cdae_save)."""
        _lib.check(self._L.cdae_save(self._h, str(path).encode()))

    def load(self, path):
        _lib.check(self._L.cdae_load(self._h, str(path).encode()))
        self._topk = 0

    def profile(self, enable=True):
        _lib.check(self._L.cdae_profile(self._h, int(enable)))

    def profile_get(self):
        """{class: (ms, launches)} since profile(True)."""
        n = len(_lib.KERNEL_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        _lib.check(self._L.cdae_profile_get(self._h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(_lib.KERNEL_CLASSES)}

    def probe_l2(self, rows, mode, row_visits, reps=20):
        """(GB/s, ms per launch) of the L2 row-load (mode 1) / row-reduction (2) / both (3) pattern."""
        g, ms = C.c_double(), C.c_double()
        _lib.check(self._L.cdae_probe_l2(self._h, rows, mode, row_visits, reps, C.byref(g), C.byref(ms)))
        return g.value, ms.value

    def pinned_array(self, src):
        """Copy of `src` in pinned host memory (cdae_host_alloc); freed with the model."""
        src = np.ascontiguousarray(src)
        p = C.c_void_p()
        _lib.check(self._L.cdae_host_alloc(C.byref(p), src.nbytes))
        buf = (C.c_char * max(src.nbytes, 1)).from_address(p.value)
        a = np.frombuffer(buf, dtype=src.dtype, count=src.size).reshape(src.shape)
        a[...] = src
        self._pinned = getattr(self, "_pinned", []) + [p]
        return a

    def synchronize(self):
        _lib.check(self._L.cdae_synchronize(self._h))


class CDAEGroup(CDAE):
    """`class CDAE` on several GPUs of ONE process (cdae_group_*, csrc/group.inl): one engine handle per
    device, users sharded like a process group, one worker thread per GPU inside every call.
    config.batch_users is the GLOBAL minibatch (0 -> 16384 per GPU)."""

    def __init__(self, config=None, devices=(0, 1)):
        super().__init__(config)
        self.devices = list(devices)
        self._g = C.c_void_p()

    def reset(self, U, I, row_ptr, col):
        self.close()
        self.U, self.I = int(U), int(I)
        self.row_ptr = _arr(row_ptr, np.int64)
        self.col = _arr(col, np.int32)
        c = self.config.to_c()
        dev = _arr(self.devices, np.int32)
        _lib.check(self._L.cdae_group_create(C.byref(c), self.U, self.I, _ptr(self.row_ptr, _lib.i64p), _ptr(self.col, _lib.i32p),
                                             _ptr(dev, _lib.i32p), len(dev), C.byref(self._g)))
        h = C.c_void_p()
        _lib.check(self._L.cdae_group_handle(self._g, 0, C.byref(h)))
        self._h = h                      # GPU 0's handle: shapes, replicated item-side reads, profiling
        return self

    def close(self):
        if getattr(self, "_g", None):
            self._L.cdae_group_destroy(self._g)
            self._g = C.c_void_p()
        self._h = C.c_void_p()

    def init_params(self, seed):
        self._topk = 0
        _lib.check(self._L.cdae_group_init_params(self._g, seed))

    def set_params(self, params):
        self._topk = 0
        for k, v in params.items():
            r, c = self.param_shape(k)
            if r * c == 0:
                continue
            a = _arr(np.asarray(v, np.float64).reshape(-1), np.float64)
            _lib.check(self._L.cdae_group_set_param(self._g, PARAM_ID[k], _ptr(a, _lib.f64p), a.size))

    def get_param(self, name):
        r, c = self.param_shape(name)
        a = np.zeros(r * c)
        if a.size:
            _lib.check(self._L.cdae_group_get_param(self._g, PARAM_ID[name], _ptr(a, _lib.f64p), a.size))
        return a.reshape(r, c) if c > 1 else a

    def train_one_iteration(self, seed=0, epoch=0, csr=None):
        st = EpochStats()
        if csr is None:
            _lib.check(self._L.cdae_group_train_epoch(self._g, seed, epoch, C.byref(st)))
        else:
            rp, cl = _arr(csr[0], np.int64), _arr(csr[1], np.int32)
            _lib.check(self._L.cdae_group_train_epoch_csr(self._g, _ptr(rp, _lib.i64p), _ptr(cl, _lib.i32p), seed, epoch, C.byref(st)))
        self.last_stats = st
        self._topk = 0
        return st

    def train_users(self, uids, keep_mask, negatives):
        u, k = _arr(uids, np.int64), _arr(keep_mask, np.uint8)
        n = None if negatives is None else _arr(negatives, np.int32)
        st = EpochStats()
        _lib.check(self._L.cdae_group_train_users(self._g, _ptr(u, _lib.i64p), len(u), _ptr(k, _lib.u8p),
                                                  None if n is None else _ptr(n, _lib.i32p), C.byref(st)))
        self._topk = 0
        return st

    def encode(self, uids, keep_mask=None, scale=1.0):
        u = _arr(uids, np.int64)
        z = np.zeros((len(u), self.config.num_dim), np.float32)
        k = None if keep_mask is None else _arr(keep_mask, np.uint8)
        _lib.check(self._L.cdae_group_encode(self._g, _ptr(u, _lib.i64p), len(u), None if k is None else _ptr(k, _lib.u8p),
                                             scale, _ptr(z, _lib.f32p)))
        return z

    def data_loss(self, seed=0):
        v = C.c_double()
        _lib.check(self._L.cdae_group_data_loss(self._g, seed, C.byref(v)))
        return v.value

    def penalty_loss(self):
        v = C.c_double()
        _lib.check(self._L.cdae_group_penalty_loss(self._g, C.byref(v)))
        return v.value

    def pre_recommend(self, topk=10):
        _lib.check(self._L.cdae_group_topn_build(self._g, topk))
        self._topk = topk

    def recommend(self, uid, topk=10):
        if self._topk != topk:
            self.pre_recommend(topk)
        ids, sc = np.zeros(topk, np.int64), np.zeros(topk, np.float32)
        _lib.check(self._L.cdae_group_topn_lookup(self._g, uid, _ptr(ids, _lib.i64p), _ptr(sc, _lib.f32p)))
        return ids, sc

    def recommend_all(self, topk=10):
        ids, sc = np.zeros((self.U, topk), np.int64), np.zeros((self.U, topk), np.float32)
        for u in range(self.U):
            ids[u], sc[u] = self.recommend(u, topk)
        return ids, sc

    def save(self, path):
        _lib.check(self._L.cdae_group_save(self._g, str(path).encode()))

    def load(self, path):
        _lib.check(self._L.cdae_group_load(self._g, str(path).encode()))
        self._topk = 0
