/* oracle/cdae_oracle.h — TEST INFRASTRUCTURE (the parity checker), NOT product code.
 *
 * Plain-C, fp64 restatement of the reference's CDAE hot path
 * (jasonyaw/CDAE @ 9b53519, /root/reference/src/model/recsys/cdae.hpp and the
 * helpers it calls).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product
 * (cdae_b200/, include/) never does.
 *
 * Pinning: the reference's own tests never touch CDAE, so
 * there are no upstream golden vectors for this path.  The oracle is pinned
 * instead against the reference ITSELF: oracle/ref_driver.cpp compiles the
 * verbatim reference headers (with stand-ins for the absent Eigen/Boost/glog/
 * gflags) into oracle/_ref/libcdae_ref.so, tests/test_oracle_vs_reference.py
 * checks this restatement against it to <=1e-12, and tests/golden/ (npz files) holds
 * vectors generated from it by tests/golden/make_golden.py.
 *
 * All index types are int64 (items/users) at this API; CSR columns are int32.
 */
#ifndef CDAE_ORACLE_H_
#define CDAE_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* libcf::LossType numbering, src/model/loss.hpp:10-18 */
enum { ORC_SQUARE = 0, ORC_LOGISTIC = 1, ORC_LOG = 2, ORC_HINGE = 3, ORC_SQUARED_HINGE = 4,
       ORC_CROSS_ENTROPY = 5, ORC_LOGM = 6 };

/* parameter blocks, same numbering as include/cdae_b200.h and ref_driver.cpp */
enum { ORC_W = 0, ORC_V, ORC_WU, ORC_B, ORC_BPRIME, ORC_UU,
       ORC_W_AG, ORC_V_AG, ORC_WU_AG, ORC_B_AG, ORC_BPRIME_AG, ORC_UU_AG, ORC_NUM_PARAMS };

/* CDAEConfig, cdae.hpp:13-31 (penalty type is always L2 on the gradient side) */
typedef struct {
  double lambda, learn_rate, corruption_ratio, beta;
  int32_t loss_type, num_dim, num_neg, num_corruptions;
  int32_t using_adagrad, asymmetric, user_factor, linear, scaled, linear_function, tanh_act;
} orc_config;

typedef struct orc_model orc_model;

/* Copies the CSR (rows must hold each user's train items; order inside a row is kept and
 * is the order positives are visited in unless an explicit order is passed).
 * Parameters start as: W,V,Wu = 0 (set them with orc_param), Uu = 1, accumulators 1e-4,
 * b = b' = 0  (cdae.hpp:109-134 minus the rand()-driven draw). */
orc_model* orc_create(const orc_config* cfg, int64_t U, int64_t I,
                      const int64_t* row_ptr, const int32_t* col);
void orc_destroy(orc_model* m);
/* Direct pointer to a parameter block (row-major rows x cols doubles). */
double* orc_param(orc_model* m, int which, int64_t* rows, int64_t* cols);
/* U[-1,1]*4*sqrt(6/(I+K)) init from the Philox stream the CUDA path uses (same values). */
void orc_init_params(orc_model* m, uint64_t seed);

/* loss.hpp gradient / evaluate; returns NaN for LOGISTIC outside (0,1) (reference CHECK-aborts) */
double orc_loss_gradient(int loss_type, double pred, double truth);
double orc_loss_evaluate(int loss_type, double pred, double truth);

/* cdae.hpp:373-416 */
void orc_hidden(const orc_model* m, int64_t uid, const int64_t* items, int64_t n, double scale,
                double* z_out);
/* cdae.hpp:418-426 */
double orc_output(const orc_model* m, const double* z, int64_t item);

/* cdae.hpp:198-358, exact order of updates.  out set = CSR row of uid, visited in
 * out_order (n_u ids) when non-NULL else CSR order; negs in draw order. */
void orc_step_sequential(orc_model* m, int64_t uid, const int64_t* in_items, int64_t n_in,
                         const int64_t* negs, int64_t n_negs, const int64_t* out_order);

/* Frozen-batch step: gradients of all users at the frozen
 * parameters, one lambda*theta per occurrence, summed, then ONE upd per touched row.
 * in_ptr / neg_ptr are CSR-style offsets (n_users+1) into in_items / negs.
 * loss_sum_out (nullable) receives sum over all outputs of loss(y, t). */
void orc_step_frozen(orc_model* m, int64_t n_users, const int64_t* uids,
                     const int64_t* in_ptr, const int64_t* in_items,
                     const int64_t* neg_ptr, const int64_t* negs, double* loss_sum_out);

/* Data-parallel decomposition of the frozen-batch step (what each GPU rank computes, SURVEY.md
 * §8e): orc_shard_gradients evaluates a SHARD of a minibatch at the frozen parameters, ADDS its
 * item-side gradients to dense_grad = [gW (I*K) | gV (I*K, asymmetric only) | gb' (I) | gb (K)]
 * and updates only the shard's user-private rows (Wu, Uu); orc_apply_dense applies the summed
 * gradient once.  shard_gradients over all shards + apply_dense == orc_step_frozen. */
void orc_shard_gradients(orc_model* m, int64_t n_users, const int64_t* uids,
                         const int64_t* in_ptr, const int64_t* in_items,
                         const int64_t* neg_ptr, const int64_t* negs, double* loss_sum_out,
                         double* dense_grad);
void orc_apply_dense(orc_model* m, const double* dense_grad, int any_steps);

/* H12, full-item-decode training (an EXTENSION: the reference has no such function):
 * the frozen-batch step above with every user's output set = all I items (target 1 on the train
 * row, 0 elsewhere) — identical to orc_step_frozen called with "negatives = every non-positive
 * item once" (tests/test_oracle_golden.py asserts that).  rounding = 1 additionally restates the
 * bf16 operand rounding of the tensor-core path (z, W', g to bf16 where they enter the three
 * contractions; b' as a bf16 hi+lo pair) so that the CUDA kernels can be checked tightly;
 * rounding = 0 is plain fp64. */
void orc_step_frozen_full(orc_model* m, int64_t n_users, const int64_t* uids, const int64_t* in_ptr,
                          const int64_t* in_items, int rounding, double* loss_sum_out);
void orc_shard_gradients_full(orc_model* m, int64_t n_users, const int64_t* uids, const int64_t* in_ptr,
                              const int64_t* in_items, int rounding, double* loss_sum_out,
                              double* dense_grad);
double orc_train_epoch_full(orc_model* m, uint64_t seed, int64_t epoch, int64_t batch_users,
                            int64_t u0, int64_t u1, int rounding);

/* cdae.hpp:162-196 with rated = the user's train row.  ids sorted by score desc
 * (exact-score ties: lower id first; the reference leaves tie order unspecified). */
int orc_recommend(const orc_model* m, int64_t uid, int64_t topk, int64_t* ids_out,
                  double* scores_out);

/* cdae.hpp:78-101 with an explicit keep mask over the CSR slots (nnz bytes, one corruption). */
double orc_data_loss(const orc_model* m, const uint8_t* keep);
/* cdae.hpp:103-107 */
double orc_penalty_loss(const orc_model* m);
/* cdae.hpp:148-159 */
void orc_user_representations(const orc_model* m, double* out);

/* evaluation.hpp:183-219 */
void orc_evaluate_rec_list(const int64_t* list, int64_t n_list, const int64_t* test_items,
                           int64_t n_test, double* out8);
/* evaluation.hpp:113-181 over a test CSR (same U); returns #users with test items */
int64_t orc_topn_evaluate(const orc_model* m, const int64_t* test_row_ptr, const int32_t* test_col,
                          double* out8);

/* ---- counter-based sampling shared (as a specification) with the CUDA path ----
 * Philox4x32-10, key = (seed lo, seed hi).
 *  keep mask of slot s of user u in pass p: word (s&3) of philox(ctr = {u, s>>2, p, 0});
 *      keep iff word > floor(q*2^32)  (q<=0 keeps all, q>=1 keeps none)
 *  negative draw d of user u in pass p: word (d&3) of philox(ctr = {u, d>>2, p, 1});
 *      item = (word * I) >> 32; while the item is in row u retry a = 1,2,..: word ((a-1)&3) of
 *      philox(ctr = {u, d, p, 2 + ((a-1)>>2)}).
 *  pass p = epoch * num_corruptions + corruption index. */
void orc_philox4x32(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t out[4]);
void orc_sample_keep(const orc_model* m, uint64_t seed, uint32_t pass, int64_t uid, uint8_t* keep_out);
void orc_sample_negatives(const orc_model* m, uint64_t seed, uint32_t pass, int64_t uid,
                          int64_t* negs_out /* n_u * num_neg */);

/* One epoch (cdae.hpp:136-146) with the Philox masks/negatives above.
 *  batch_users <= 1 : sequential reference semantics (the CPU baseline "port");
 *  batch_users  > 1 : frozen-batch of that many consecutive users.
 * Users [u0,u1).  Returns the sum over outputs of loss(y,t) seen during the pass
 * (frozen mode) or 0 (sequential mode). */
double orc_train_epoch(orc_model* m, uint64_t seed, int64_t epoch, int64_t batch_users,
                       int64_t u0, int64_t u1);
/* Our multi-core CPU variant (NOT the reference's): Hogwild user-parallel sequential steps. */
double orc_train_epoch_hogwild(orc_model* m, uint64_t seed, int64_t epoch, int64_t u0, int64_t u1,
                               int n_threads);

#ifdef __cplusplus
}
#endif
#endif
