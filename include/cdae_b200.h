/* cdae_b200.h — C ABI of the B200-native CDAE training / scoring engine.
 *
 * This is the drop-in boundary for the hot path of jasonyaw/CDAE (libcf), i.e. everything
 * `class CDAE` (src/model/recsys/cdae.hpp) computes.  The reference has no FFI for this path
 * (it is a header-only C++ class that Solver<Model> / TOPN_Evaluation<Model> duck-type
 * against), so each entry point below names the reference member function
 * whose work it takes over; the host-side `libcf::CDAE` in cdae_b200/host/ keeps the
 * reference's class surface and is a thin caller of these functions.
 *
 * Conventions
 *  - plain C: opaque handle, pointers + sizes, no C++ / torch types.
 *  - every function returns 0 on success, a negative CDAE_E_* code otherwise;
 *    cdae_last_error() returns a thread-local message for the last failure.  (The reference's
 *    convention is glog CHECK -> abort, cdae.hpp:83,139,187; the host class maps a non-zero
 *    return to LOG(FATAL) to keep that behaviour.)
 *  - all host buffers are caller-owned and may be pageable or pinned (cdae_host_alloc gives
 *    pinned memory); device memory is owned by the handle.
 *  - one handle is driven by one host thread at a time, except cdae_topn_lookup, which only
 *    reads the table cdae_topn_build produced and is safe to call concurrently (the reference
 *    calls recommend() from ThreadPool workers, evaluation.hpp:137-158).
 *  - parameters live on the device in fp32 (the reference keeps fp64 Eigen matrices,
 *    base/mat.hpp:12,19-22); the boundary exchanges them as row-major doubles, the
 *    reference's own type.
 *  - CSR rows must be strictly ascending (needed for the on-device negative sampler);
 *    every user that is trained must have >= 1 item (CHECK at cdae.hpp:139).
 */
#ifndef CDAE_B200_H_
#define CDAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDAE_B200_ABI_VERSION 1

/* error codes */
#define CDAE_OK 0
#define CDAE_E_INVALID (-1)     /* bad argument / shape / unsupported configuration */
#define CDAE_E_CUDA (-2)        /* CUDA runtime error (message has the CUDA string) */
#define CDAE_E_NCCL (-3)        /* NCCL error or NCCL not loadable */
#define CDAE_E_STATE (-4)       /* call order (e.g. topn_lookup before topn_build) */
#define CDAE_E_NUMERIC (-5)     /* LOGISTIC loss fed a score outside (0,1): reference aborts (loss.hpp:96) */

/* libcf::LossType, src/model/loss.hpp:10-18 */
enum cdae_loss {
  CDAE_LOSS_SQUARE = 0, CDAE_LOSS_LOGISTIC = 1, CDAE_LOSS_LOG = 2, CDAE_LOSS_HINGE = 3,
  CDAE_LOSS_SQUARED_HINGE = 4, CDAE_LOSS_CROSS_ENTROPY = 5, CDAE_LOSS_LOGM = 6
};

/* parameter blocks of class CDAE (cdae.hpp:430-439); *_AG are the AdaGrad accumulators */
enum cdae_param {
  CDAE_P_W = 0,     /* I x K  encoder (and, tied, decoder) item weights            */
  CDAE_P_V,         /* I x K  decoder item weights, only when asymmetric           */
  CDAE_P_WU,        /* U x K  per-user input embedding, only when user_factor      */
  CDAE_P_B,         /* K      hidden bias                                          */
  CDAE_P_BPRIME,    /* I      output bias                                          */
  CDAE_P_UU,        /* U x K  per-user multiplicative map, only when linear_function */
  CDAE_P_W_AG, CDAE_P_V_AG, CDAE_P_WU_AG, CDAE_P_B_AG, CDAE_P_BPRIME_AG, CDAE_P_UU_AG,
  CDAE_P_COUNT
};

/* libcf::CDAEConfig (cdae.hpp:13-31) + device options.  Fill with cdae_config_default()
 * first; fields added by later ABI versions then keep their defaults. */
typedef struct cdae_config {
  double lambda;              /* L2 coefficient, applied per touched row per occurrence */
  double learn_rate;
  double corruption_ratio;    /* q: an input item is kept iff uniform > q (cdae.hpp:366) */
  double beta;                /* AdaGrad denominator offset: g / (beta + sqrt(acc))      */
  int32_t loss_type;          /* enum cdae_loss */
  int32_t num_dim;            /* K */
  int32_t num_neg;            /* negatives per positive */
  int32_t num_corruptions;
  int32_t using_adagrad;      /* 0: plain SGD */
  int32_t asymmetric;         /* 0: tied weights (decoder = W), 1: separate V */
  int32_t user_factor;
  int32_t linear;             /* identity activation */
  int32_t scaled;             /* scale inputs by 1/(1-q) */
  int32_t linear_function;
  int32_t tanh_act;
  /* ---- device options (no reference counterpart) ---- */
  int32_t batch_users;        /* users per frozen minibatch; 0 -> 16384.  1 reproduces the
                                 reference's per-user online step. */
  int32_t device;             /* CUDA device ordinal */
  int32_t full_decode;        /* 1: full-item decode — every user's output set is ALL items (target 1
                                 on the train row, 0 elsewhere) instead of positives + num_neg
                                 sampled negatives; the K x I contraction runs on tcgen05/TMEM with
                                 bf16 operands (no reference function: SURVEY.md H12).  Needs
                                 CROSS_ENTROPY or SQUARE loss and num_dim <= 256; batch_users 0 ->
                                 128 x SM count. */
  int32_t reserved[6];
} cdae_config_t;

typedef struct cdae_epoch_stats {
  int64_t user_steps;         /* (user, corruption) pairs trained                     */
  int64_t outputs;            /* positives + negatives scored                          */
  int64_t inputs_kept;        /* input items that survived corruption                  */
  double loss_sum;            /* sum over scored outputs of loss(y, t); 0 with full_decode (the tensor-core
                                 epilogue only forms the loss GRADIENT; use cdae_data_loss) */
  double device_ms;           /* device time of the call, CUDA events on the handle's stream */
  int64_t kernel_launches;    /* kernels this call launched                            */
  int64_t h2d_bytes, d2h_bytes; /* bytes this call copied across PCIe                  */
} cdae_epoch_stats_t;

typedef struct cdae_handle cdae_handle;

int cdae_abi_version(void);
const char* cdae_last_error(void);

/* CDAEConfig() defaults (cdae.hpp:14-30): lambda .01, lr .1, LOGISTIC, K 10, adagrad, q .5,
 * cnum 1, tied, user_factor, sigmoid, num_neg 5, scaled, beta 0. */
int cdae_config_default(cdae_config_t* cfg);

/* CDAE::CDAE + CDAE::reset (cdae.hpp:39-74,109-134; RecsysModelBase::reset,
 * recsys_model_base.hpp:29-34): takes the user->items structure the reference builds with
 * Data::get_feature_pair_label_hashtable(0,1) as CSR, uploads it, allocates parameters
 * (accumulators 1e-4, b = b' = 0, Uu = 1; W, V, Wu zero until cdae_init_params / cdae_set_param). */
int cdae_create(const cdae_config_t* cfg, int64_t num_users, int64_t num_items,
                const int64_t* row_ptr, const int32_t* col_idx, cdae_handle** out);
int cdae_destroy(cdae_handle* h);

/* The Random(I,K) * 4*sqrt(6/(I+K)) draw of CDAE::reset (cdae.hpp:112-121), from a
 * counter-based Philox stream instead of rand() (values documented in DESIGN.md). */
int cdae_init_params(cdae_handle* h, uint64_t seed);

/* No reference counterpart (parameters are private there, cdae.hpp:428): read / write one
 * block as row-major doubles.  n must equal rows*cols of the block; absent blocks have n = 0. */
int cdae_param_shape(cdae_handle* h, int which, int64_t* rows, int64_t* cols);
int cdae_set_param(cdae_handle* h, int which, const double* src, int64_t n);
int cdae_get_param(cdae_handle* h, int which, double* dst, int64_t n);
/* n selected rows of a table (dst: n x K doubles) or n selected entries of b / b' (dst: n
 * doubles) — what CDAE::get_output_values (cdae.hpp:418-426) reads: W'.row(i) and b'(i). */
int cdae_get_param_rows(cdae_handle* h, int which, const int64_t* rows, int64_t n, double* dst);

/* CDAE::train_one_iteration (cdae.hpp:136-146): one pass over users [0,U) in minibatches of
 * batch_users, masks and negatives from Philox(seed, epoch) (spec in DESIGN.md; identical
 * on any GPU count).  In a process group (cdae_dist_init) each rank trains its shard of every
 * minibatch and the dense item-side gradients are all-reduced once per minibatch. */
int cdae_train_epoch(cdae_handle* h, uint64_t seed, int64_t epoch, cdae_epoch_stats_t* stats);

/* Same, with the training CSR taken from HOST memory on every call, the way
 * train_one_iteration(const Data&) receives its data each epoch.  Shapes must match
 * cdae_create (same U; nnz may differ).  H2D copy and the D2H read of stats are inside
 * the call.  The upload runs on a second stream, one piece per minibatch, under the kernels of the
 * previous minibatch.  row_ptr is checked on the host when it changed; col is checked ON THE DEVICE as
 * each minibatch's rows arrive (ids in [0, I), rows strictly ascending; in a process group each rank
 * checks the rows it trains): a violation returns CDAE_E_INVALID — the offending minibatch and every
 * later one leave the parameters untouched, earlier minibatches of the call have been trained
 * (out-of-range ids are clamped in the device copy, so no kernel indexes outside a table) — and
 * training calls are refused until a valid CSR is passed. */
int cdae_train_epoch_csr(cdae_handle* h, const int64_t* row_ptr, const int32_t* col_idx,
                         uint64_t seed, int64_t epoch, cdae_epoch_stats_t* stats);

/* CDAE::train_one_user_corruption (cdae.hpp:198-358) for n DISTINCT users as ONE frozen
 * minibatch with EXPLICIT randomness: keep_mask has one byte per train item of each listed
 * user (CSR order, concatenated); negatives has n_u*num_neg item ids per user, concatenated
 * (each must be outside that user's row; ignored, may be NULL, with full_decode).  n = 1 is the
 * reference's online step. */
int cdae_train_users(cdae_handle* h, const int64_t* uids, int64_t n, const uint8_t* keep_mask,
                     const int32_t* negatives, cdae_epoch_stats_t* stats);

/* CDAE::get_hidden_values (cdae.hpp:373-416) for n users: z_out is n x K floats.
 * keep_mask NULL -> uncorrupted input; scale multiplies the summed rows (the reference
 * passes 1/(1-q) when scaled, 1 otherwise).  Covers get_user_representations (cdae.hpp:148). */
int cdae_encode(cdae_handle* h, const int64_t* uids, int64_t n, const uint8_t* keep_mask,
                double scale, float* z_out);

/* CDAE::data_loss (cdae.hpp:78-101): fresh Philox corruption, sum over users of the loss of
 * their positives, averaged over num_corruptions.  CDAE::penalty_loss (cdae.hpp:103-107). */
int cdae_data_loss(cdae_handle* h, uint64_t seed, double* out);
int cdae_penalty_loss(cdae_handle* h, double* out);

/* CDAE::recommend (cdae.hpp:162-196) for ALL users at once — the batch point is the
 * pre_recommend() hook (recsys_model_base.hpp:72, evaluation.hpp:135).  Scores every item
 * against every user's uncorrupted hidden vector, skips the user's train items, keeps the
 * top-k by the reference's rule (strict improvement, ties keep the lower id), ids sorted by
 * score descending.  cdae_topn_lookup copies one user's list (thread-safe). */
int cdae_topn_build(cdae_handle* h, int32_t topk);
int cdae_topn_lookup(cdae_handle* h, int64_t uid, int64_t* ids_out, float* scores_out);
/* The candidate phase of the last cdae_topn_build: path 1 = bf16 tcgen05/TMEM contraction whose
 * lists are PROVEN exact per user by an error bound (verified_users), the rest (redone_users)
 * recomputed by the exact fp32 kernel; path 0 = fp32 kernel for everyone (K > 318, topk > 16, or
 * CDAE_B200_TOPN=fp32 in the environment). */
int cdae_topn_stats(cdae_handle* h, int32_t* path, int64_t* verified_users, int64_t* redone_users);
/* Size of the probe table of the last cdae_topn_build (tensor path): the first sweep started from per-user
 * thresholds derived from the items with the largest mean-user score; 0 = no probe pass (fewer than 512
 * items, the fp32 path, or switched off with CDAE_B200_TOPN_PROBE=0).  Only the cost of the build depends on
 * it — every list is verified or recomputed exactly either way. */
int cdae_topn_probe_items(cdae_handle* h, int32_t* items_out);
/* Whole table, U x topk (ids) and U x topk (scores; nullable). */
int cdae_topn_fetch(cdae_handle* h, int64_t* ids_out, float* scores_out);
/* TOPN_Evaluation::evaluate (evaluation.hpp:113-181) on the built table against a test CSR:
 * out8 = P@1,P@5,P@10,R@1,R@5,R@10,MAP@5,MAP@10 averaged over users with test items. */
int cdae_topn_evaluate(cdae_handle* h, const int64_t* test_row_ptr, const int32_t* test_col,
                       double* out8, int64_t* users_evaluated);

/* ---- the data path in front of the hot path (host only) ----
 * Data::load(file, RECSYS, parser, skip_header) (data-inl.hpp:45-64) for "user<delim>item" lines as
 * apps/yelp parses them (yelp.cpp:60-66; split_line drops empty tokens, file_utils.hpp:15-25; empty
 * lines are skipped and not counted, file_line_reader-inl.hpp:12-19), with the reference's id
 * assignment — dense ids in FIRST-SEEN order per column (FeatureGroupInfo::get_index,
 * instance-inl.hpp:22-37) — straight into the CSR cdae_create() takes (rows ascending, duplicate
 * pairs collapsed like the hash of hashes of recsys_model_base.hpp:29-34 does).  A line that does
 * not have exactly two fields is CDAE_E_INVALID (the reference CHECK-aborts, yelp.cpp:62). */
typedef struct cdae_dataset cdae_dataset;
int cdae_dataset_load_pairs(const char* path, const char* delimiters /* NULL -> " " */,
                            int32_t skip_header, cdae_dataset** out);
int cdae_dataset_info(const cdae_dataset* d, int64_t* users, int64_t* items, int64_t* instances);
/* Data::random_split_by_feature_group(train, test, 0, test_ratio) (data-inl.hpp:231-272): per user,
 * a uniformly random floor(n_u * test_ratio) of its instances go to test.  Philox stream keyed by
 * (seed, user) instead of the reference's time-seeded mt19937_64. */
int cdae_dataset_split(cdae_dataset* d, double test_ratio, uint64_t seed);
/* which: 0 = all pairs, 1 = train, 2 = test (1, 2 need cdae_dataset_split).  row_ptr: users + 1. */
int cdae_dataset_nnz(const cdae_dataset* d, int32_t which, int64_t* nnz);
int cdae_dataset_csr(const cdae_dataset* d, int32_t which, int64_t* row_ptr, int32_t* col_idx);
/* the raw string of a dense id (group 0 = users, 1 = items): FeatureGroupInfo::raw_str_map_ */
int cdae_dataset_raw_id(const cdae_dataset* d, int32_t group, int64_t idx, const char** out);
int cdae_dataset_free(cdae_dataset* d);
/* Data set cache: the counterpart of Data::save / Data::load (data.hpp:25-33, 52-60;
 * io/serialize.hpp:16-46 — gzip over a boost binary archive, a byte format that cannot be reproduced without
 * Boost).  Own versioned binary format: raw id tables, the instances in file order and, if cdae_dataset_split
 * has run, the train / test CSRs — so that a CPU baseline run and a GPU run share one split.  cdae_dataset_load
 * checks every size against the file and the CSR invariants cdae_create relies on; a file that fails is
 * CDAE_E_INVALID. */
int cdae_dataset_save(const cdae_dataset* d, const char* path);
int cdae_dataset_load(const char* path, cdae_dataset** out);

/* Model checkpoint (the reference has none — its save/load only cover Data):
 * versioned binary file with the config, the shape and every parameter block incl. AdaGrad state
 * as doubles; the config is stored field by field at fixed widths.  cdae_load needs a handle created
 * with the same shape and the same structural options (asymmetric, user_factor, linear_function); the
 * handle's other hyper-parameters stay in force (a difference is noted in cdae_last_error(), rc 0).
 * In a process group cdae_save is COLLECTIVE (user-private blocks are assembled from their owning
 * ranks): every rank calls it, only rank 0 writes `path`; every rank calls cdae_load on the same file. */
int cdae_save(cdae_handle* h, const char* path);
int cdae_load(cdae_handle* h, const char* path);

/* Data parallelism over the GPUs of one node: one process per GPU, each owning the user
 * shard [rank*U/world, (rank+1)*U/world) of every minibatch; item-side parameters are
 * replicated and kept identical by all-reducing the dense gradients.  nccl_unique_id is
 * the 128-byte ncclUniqueId created on rank 0 (cdae_dist_unique_id) and distributed by the
 * caller (torch.distributed / MPI / a file). */
int cdae_dist_unique_id(void* id128_out);
int cdae_dist_init(cdae_handle* h, int32_t rank, int32_t world, const void* nccl_unique_id);

/* Peer-memory mode of the per-minibatch combine step (csrc/p2p_allreduce.cuh): instead of an NCCL
 * all-reduce followed by the same dense optimiser pass on every rank, ONE kernel per rank reduce-scatters
 * the gradients with loads from NVLink peer memory, applies the optimiser step to the rank's 1/G slice
 * (AdaGrad state is sharded: a rank keeps only its slice current) and all-gathers the updated
 * parameters with peer stores.  Every rank exports CUDA IPC handles of its gradient buffers, its
 * item-side parameter buffer and its flag array (a 256-byte record), the caller gathers the records in
 * rank order (world x 256 bytes) and hands the table to every rank.  Needs cdae_dist_init first, 2..8
 * ranks on one node with peer access; every rank must open before the next training call.  NCCL remains
 * in use for everything else (parameter read-back, data_loss).  cdae_get_param of an accumulator block
 * assembles the slices (collective). */
int cdae_dist_p2p_export(cdae_handle* h, void* record256_out);
int cdae_dist_p2p_open(cdae_handle* h, const void* all_handles);

/* NVLS mode of the same fused combine step (csrc/mc_nvls.inl, p2p::mc_step_kernel): the item-side
 * buffers of every rank are bound to ONE NVSwitch multicast object; the kernel sums a rank's slice of the
 * gradients with multimem.ld_reduce (reduced inside the switch: 1/G of the buffer inbound per GPU instead
 * of (G-1)/G), applies its slice of the optimiser step and publishes the updated parameters with
 * multimem.st.  Use INSTEAD of cdae_dist_p2p_export / _open, after cdae_dist_init:
 *   rank 0:     cdae_dist_mc_create  -> a POSIX file descriptor of the multicast object; the caller passes
 *               it to the other ranks (SCM_RIGHTS over a unix socket; cdae_b200/model.py has a helper)
 *   every rank: cdae_dist_mc_attach(fd)   (rank 0 passes the descriptor it created, or -1)
 *   -- caller's barrier: every rank has attached --
 *   every rank: cdae_dist_mc_bind          (allocates, binds and maps the rank's block, moves the item side)
 *   -- caller's barrier --                 then train.
 * Each returns CDAE_E_STATE where multicast is unavailable (no NVSwitch, driver without NVLS, device
 * attribute MULTICAST_SUPPORTED = 0): fall back to cdae_dist_p2p_* or plain NCCL. */
int cdae_dist_mc_create(cdae_handle* h, int32_t* fd_out);
int cdae_dist_mc_attach(cdae_handle* h, int32_t fd);
int cdae_dist_mc_bind(cdae_handle* h);

/* ---- single-process multi-GPU mode (csrc/group.inl) -------------------------------------------------
 * The reference's app is one process (Solver<CDAE>::train, solver-inl.hpp:19,53,55).  A cdae_group owns
 * one engine handle per device and drives them from one worker thread per GPU inside every call, wired
 * like a process group of `n` ranks (same user sharding, same fused combine step over peer memory —
 * NVLS where available — with NCCL for the read-backs), so libcf::CDAE can use the whole box:
 * cdae_b200/host/model/recsys/cdae.hpp switches to it with CDAE_B200_GPUS=n.  cfg->batch_users is the
 * GLOBAL minibatch (0 -> 16384 per GPU); devices NULL -> 0..n-1.  Every cdae_group_X is the collective
 * form of cdae_X; cdae_group_topn_lookup (thread-safe) and cdae_group_encode route each user to the
 * GPU that trains it.  cdae_group_handle exposes one GPU's handle, e.g. for item-side (replicated) reads. */
typedef struct cdae_group cdae_group;
int cdae_group_create(const cdae_config_t* cfg, int64_t num_users, int64_t num_items, const int64_t* row_ptr,
                      const int32_t* col_idx, const int32_t* devices, int32_t n, cdae_group** out);
int cdae_group_destroy(cdae_group* g);
int cdae_group_size(cdae_group* g, int32_t* n);
int cdae_group_handle(cdae_group* g, int32_t rank, cdae_handle** out);
int cdae_group_init_params(cdae_group* g, uint64_t seed);
int cdae_group_set_param(cdae_group* g, int which, const double* src, int64_t n);
int cdae_group_get_param(cdae_group* g, int which, double* dst, int64_t n);
int cdae_group_get_param_rows(cdae_group* g, int which, const int64_t* rows, int64_t n, double* dst);
int cdae_group_train_epoch(cdae_group* g, uint64_t seed, int64_t epoch, cdae_epoch_stats_t* stats);
int cdae_group_train_epoch_csr(cdae_group* g, const int64_t* row_ptr, const int32_t* col_idx, uint64_t seed,
                               int64_t epoch, cdae_epoch_stats_t* stats);
int cdae_group_train_users(cdae_group* g, const int64_t* uids, int64_t n, const uint8_t* keep_mask,
                           const int32_t* negatives, cdae_epoch_stats_t* stats);
int cdae_group_encode(cdae_group* g, const int64_t* uids, int64_t n, const uint8_t* keep_mask, double scale,
                      float* z_out);
int cdae_group_data_loss(cdae_group* g, uint64_t seed, double* out);
int cdae_group_penalty_loss(cdae_group* g, double* out);
int cdae_group_topn_build(cdae_group* g, int32_t topk);
int cdae_group_topn_lookup(cdae_group* g, int64_t uid, int64_t* ids_out, float* scores_out);
int cdae_group_save(cdae_group* g, const char* path);
int cdae_group_load(cdae_group* g, const char* path);

/* Per-kernel-class device timing for benchmarks: when enabled, every kernel launch is
 * bracketed by CUDA events on the handle's stream; cdae_profile_get returns, per class,
 * the summed milliseconds and the number of launches since cdae_profile(h, 1). */
enum cdae_kernel_class {
  CDAE_K_SAMPLE = 0, CDAE_K_GATHER, CDAE_K_ACTIVATE, CDAE_K_DECODE, CDAE_K_HIDDEN_BWD,
  CDAE_K_SCATTER, CDAE_K_ALLREDUCE, CDAE_K_APPLY, CDAE_K_TOPN /* tcgen05 candidate kernel */,
  CDAE_K_TOPN_PACK, CDAE_K_TOPN_RERANK, CDAE_K_TOPN_EXACT /* fp32 candidate kernel */,
  /* full-item decode training (tcgen05): scores + loss gradient, hidden gradient, item gradient */
  CDAE_K_FD_PACK, CDAE_K_FD_SCORE, CDAE_K_FD_HIDDEN, CDAE_K_FD_ITEMGRAD, CDAE_K_COUNT
};
int cdae_profile(cdae_handle* h, int32_t enable);
int cdae_profile_get(cdae_handle* h, double* ms_out /*[CDAE_K_COUNT]*/,
                     int64_t* launches_out /*[CDAE_K_COUNT]*/);

/* Measurement aid (no reference counterpart): the L2 roofline of the sampled decode's access pattern.
 * When the item tables and their gradients fit the 126 MB L2 (config B: 38 MB) the decode kernel is
 * bound by L2 transactions, not HBM.  This runs a kernel that issues only those transactions —
 * `row_visits` uniformly random rows of a `rows` x ld fp32 table, 16-byte vector loads (mode & 1)
 * and / or 16-byte vector reductions into a second table (mode & 2), in the handle's row geometry —
 * and returns bytes moved per second (GB/s) and the average launch time over `reps` launches. */
int cdae_probe_l2(cdae_handle* h, int64_t rows, int32_t mode, int64_t row_visits, int32_t reps,
                  double* gbs_out, double* ms_out);

/* Measurement aid: the per-minibatch combine step alone, `reps` times on a synthetic non-zero gradient
 * (collective in a process group; with cdae_profile on, the allreduce / apply classes hold its pure
 * device time — no user work in front of it, hence no rank skew).  Parameters drift by ~1e-10 per call. */
int cdae_debug_combine(cdae_handle* h, int32_t reps);
/* %globaltimer (ns) at the phase boundaries of the last fused combine kernel cdae_debug_combine launched:
 * [0] entry, [1] all peers' gradients complete, [2] block 0 done with its slice, [3] all blocks done,
 * [4] all peers' stores are in. */
int cdae_debug_combine_times(cdae_handle* h, uint64_t* out5);

/* pinned host memory for buffers that cross the boundary every step */
int cdae_host_alloc(void** ptr, int64_t bytes);
int cdae_host_free(void* ptr);

/* Blocks until all work queued on the handle's stream has finished. */
int cdae_synchronize(cdae_handle* h);
/* The handle's CUDA stream (cudaStream_t as void*), for callers that time with their own events. */
int cdae_stream(cdae_handle* h, void** stream_out);

#ifdef __cplusplus
}
#endif
#endif /* CDAE_B200_H_ */
