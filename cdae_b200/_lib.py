"""ctypes view of include/cdae_b200.h.  There is NO CPU fallback: if libcdae_b200.so is missing
or does not load, importing the bindings raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("CDAE_B200_LIB") or os.path.join(HERE, "libcdae_b200.so")   # the override is for A/B builds of the same sources

LOSS = {"SQUARE": 0, "LOGISTIC": 1, "LOG": 2, "HINGE": 3, "SQUARED_HINGE": 4, "CE": 5,
        "CROSS_ENTROPY": 5, "LOGM": 6}
PARAMS = ["W", "V", "Wu", "b", "b_prime", "Uu",
          "W_ag", "V_ag", "Wu_ag", "b_ag", "b_prime_ag", "Uu_ag"]
PARAM_ID = {n: i for i, n in enumerate(PARAMS)}

i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)


class Config(C.Structure):
    """struct cdae_config (include/cdae_b200.h)"""
    _fields_ = [("lambda_", C.c_double), ("learn_rate", C.c_double),
                ("corruption_ratio", C.c_double), ("beta", C.c_double),
                ("loss_type", C.c_int32), ("num_dim", C.c_int32), ("num_neg", C.c_int32),
                ("num_corruptions", C.c_int32), ("using_adagrad", C.c_int32),
                ("asymmetric", C.c_int32), ("user_factor", C.c_int32), ("linear", C.c_int32),
                ("scaled", C.c_int32), ("linear_function", C.c_int32), ("tanh_act", C.c_int32),
                ("batch_users", C.c_int32), ("device", C.c_int32), ("full_decode", C.c_int32),
                ("reserved", C.c_int32 * 6)]


class EpochStats(C.Structure):
    """struct cdae_epoch_stats (include/cdae_b200.h)"""
    _fields_ = [("user_steps", C.c_int64), ("outputs", C.c_int64), ("inputs_kept", C.c_int64),
                ("loss_sum", C.c_double), ("device_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/cdae_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "cdae_abi_version": (C.c_int, []),
    "cdae_last_error": (C.c_char_p, []),
    "cdae_config_default": (C.c_int, [C.POINTER(Config)]),
    "cdae_create": (C.c_int, [C.POINTER(Config), C.c_int64, C.c_int64, i64p, i32p,
                              C.POINTER(C.c_void_p)]),
    "cdae_destroy": (C.c_int, [C.c_void_p]),
    "cdae_init_params": (C.c_int, [C.c_void_p, C.c_uint64]),
    "cdae_param_shape": (C.c_int, [C.c_void_p, C.c_int, i64p, i64p]),
    "cdae_set_param": (C.c_int, [C.c_void_p, C.c_int, f64p, C.c_int64]),
    "cdae_get_param": (C.c_int, [C.c_void_p, C.c_int, f64p, C.c_int64]),
    "cdae_get_param_rows": (C.c_int, [C.c_void_p, C.c_int, i64p, C.c_int64, f64p]),
    "cdae_train_epoch": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.POINTER(EpochStats)]),
    "cdae_train_epoch_csr": (C.c_int, [C.c_void_p, i64p, i32p, C.c_uint64, C.c_int64,
                                       C.POINTER(EpochStats)]),
    "cdae_train_users": (C.c_int, [C.c_void_p, i64p, C.c_int64, u8p, i32p, C.POINTER(EpochStats)]),
    "cdae_encode": (C.c_int, [C.c_void_p, i64p, C.c_int64, u8p, C.c_double, f32p]),
    "cdae_data_loss": (C.c_int, [C.c_void_p, C.c_uint64, f64p]),
    "cdae_penalty_loss": (C.c_int, [C.c_void_p, f64p]),
    "cdae_topn_build": (C.c_int, [C.c_void_p, C.c_int32]),
    "cdae_topn_lookup": (C.c_int, [C.c_void_p, C.c_int64, i64p, f32p]),
    "cdae_topn_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), i64p, i64p]),
    "cdae_topn_probe_items": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "cdae_topn_fetch": (C.c_int, [C.c_void_p, i64p, f32p]),
    "cdae_topn_evaluate": (C.c_int, [C.c_void_p, i64p, i32p, f64p, i64p]),
    "cdae_dist_unique_id": (C.c_int, [C.c_void_p]),
    "cdae_dist_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "cdae_dist_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cdae_dist_p2p_open": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cdae_dist_mc_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "cdae_dist_mc_attach": (C.c_int, [C.c_void_p, C.c_int32]),
    "cdae_dist_mc_bind": (C.c_int, [C.c_void_p]),
    "cdae_group_create": (C.c_int, [C.POINTER(Config), C.c_int64, C.c_int64, i64p, i32p, i32p, C.c_int32, C.POINTER(C.c_void_p)]),
    "cdae_group_destroy": (C.c_int, [C.c_void_p]),
    "cdae_group_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "cdae_group_handle": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "cdae_group_init_params": (C.c_int, [C.c_void_p, C.c_uint64]),
    "cdae_group_set_param": (C.c_int, [C.c_void_p, C.c_int, f64p, C.c_int64]),
    "cdae_group_get_param": (C.c_int, [C.c_void_p, C.c_int, f64p, C.c_int64]),
    "cdae_group_get_param_rows": (C.c_int, [C.c_void_p, C.c_int, i64p, C.c_int64, f64p]),
    "cdae_group_train_epoch": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.POINTER(EpochStats)]),
    "cdae_group_train_epoch_csr": (C.c_int, [C.c_void_p, i64p, i32p, C.c_uint64, C.c_int64, C.POINTER(EpochStats)]),
    "cdae_group_train_users": (C.c_int, [C.c_void_p, i64p, C.c_int64, u8p, i32p, C.POINTER(EpochStats)]),
    "cdae_group_encode": (C.c_int, [C.c_void_p, i64p, C.c_int64, u8p, C.c_double, f32p]),
    "cdae_group_data_loss": (C.c_int, [C.c_void_p, C.c_uint64, f64p]),
    "cdae_group_penalty_loss": (C.c_int, [C.c_void_p, f64p]),
    "cdae_group_topn_build": (C.c_int, [C.c_void_p, C.c_int32]),
    "cdae_group_topn_lookup": (C.c_int, [C.c_void_p, C.c_int64, i64p, f32p]),
    "cdae_group_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cdae_group_load": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cdae_profile": (C.c_int, [C.c_void_p, C.c_int32]),
    "cdae_profile_get": (C.c_int, [C.c_void_p, f64p, i64p]),
    "cdae_probe_l2": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, f64p, f64p]),
    "cdae_debug_combine": (C.c_int, [C.c_void_p, C.c_int32]),
    "cdae_debug_combine_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "cdae_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "cdae_host_free": (C.c_int, [C.c_void_p]),
    "cdae_synchronize": (C.c_int, [C.c_void_p]),
    "cdae_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "cdae_dataset_load_pairs": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "cdae_dataset_info": (C.c_int, [C.c_void_p, i64p, i64p, i64p]),
    "cdae_dataset_split": (C.c_int, [C.c_void_p, C.c_double, C.c_uint64]),
    "cdae_dataset_nnz": (C.c_int, [C.c_void_p, C.c_int32, i64p]),
    "cdae_dataset_csr": (C.c_int, [C.c_void_p, C.c_int32, i64p, i32p]),
    "cdae_dataset_raw_id": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_char_p)]),
    "cdae_dataset_free": (C.c_int, [C.c_void_p]),
    "cdae_dataset_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cdae_dataset_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "cdae_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cdae_load": (C.c_int, [C.c_void_p, C.c_char_p]),
}

KERNEL_CLASSES = ["sample", "gather", "activate", "decode", "hidden_bwd", "scatter", "allreduce",
                  "apply", "topn", "topn_pack", "topn_rerank", "topn_exact",
                  "fd_pack", "fd_score", "fd_hidden", "fd_itemgrad"]

_lib = None


class CdaeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cdae_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads the CUDA library; raises if it is absent (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError("%s is missing - run `python -m cdae_b200.build` "
                              "(or __graft_entry__.build()); there is no CPU fallback" % SO)
        L = C.CDLL(SO)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)       # AttributeError if the symbol is not exported
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise CdaeError(rc, lib().cdae_last_error().decode("utf-8", "replace"))
