/* oracle/cdae_oracle.c — TEST INFRASTRUCTURE (the parity checker), NOT product code.
 * See cdae_oracle.h for scope and pinning.  Every function cites the lines of
 * /root/reference/src it restates.  Arithmetic is IEEE double, evaluated
 * element by element in the reference's source order (Eigen's coefficient-wise
 * semantics); build with -ffp-contract=off so no FMA contraction sneaks in. */
#include "cdae_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

struct orc_model {
  orc_config cfg;
  int64_t U, I;
  int K;
  int64_t* row_ptr;
  int32_t* col;
  int rows_sorted; /* every CSR row strictly ascending -> membership by binary search */
  double* p[ORC_NUM_PARAMS];
};

static int64_t param_rows(const orc_model* m, int which) {
  switch (which) {
    case ORC_W: case ORC_V: case ORC_W_AG: case ORC_V_AG: return m->I;
    case ORC_WU: case ORC_UU: case ORC_WU_AG: case ORC_UU_AG: return m->U;
    case ORC_B: case ORC_B_AG: return m->K;
    case ORC_BPRIME: case ORC_BPRIME_AG: return m->I;
  }
  return 0;
}
static int64_t param_cols(const orc_model* m, int which) {
  switch (which) {
    case ORC_B: case ORC_B_AG: case ORC_BPRIME: case ORC_BPRIME_AG: return 1;
    default: return m->K;
  }
}

/* cdae.hpp:109-134: accumulators Constant(1e-4); b, b' zero; Uu Constant(1). */
orc_model* orc_create(const orc_config* cfg, int64_t U, int64_t I, const int64_t* row_ptr,
                      const int32_t* col) {
  orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
  m->cfg = *cfg;
  m->U = U;
  m->I = I;
  m->K = cfg->num_dim;
  m->row_ptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(U + 1));
  memcpy(m->row_ptr, row_ptr, sizeof(int64_t) * (size_t)(U + 1));
  int64_t nnz = row_ptr[U];
  m->col = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  memcpy(m->col, col, sizeof(int32_t) * (size_t)nnz);
  m->rows_sorted = 1;
  for (int64_t u = 0; u < U && m->rows_sorted; ++u)
    for (int64_t s = row_ptr[u] + 1; s < row_ptr[u + 1]; ++s)
      if (col[s - 1] >= col[s]) {
        m->rows_sorted = 0;
        break;
      }
  for (int w = 0; w < ORC_NUM_PARAMS; ++w) {
    size_t n = (size_t)(param_rows(m, w) * param_cols(m, w));
    m->p[w] = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    double init = 0.0;
    if (w >= ORC_W_AG) init = 0.0001;
    if (w == ORC_UU) init = 1.0;
    if (init != 0.0)
      for (size_t i = 0; i < n; ++i) m->p[w][i] = init;
  }
  return m;
}

void orc_destroy(orc_model* m) {
  if (!m) return;
  for (int w = 0; w < ORC_NUM_PARAMS; ++w) free(m->p[w]);
  free(m->row_ptr);
  free(m->col);
  free(m);
}

double* orc_param(orc_model* m, int which, int64_t* rows, int64_t* cols) {
  if (which < 0 || which >= ORC_NUM_PARAMS) return NULL;
  if (rows) *rows = param_rows(m, which);
  if (cols) *cols = param_cols(m, which);
  return m->p[which];
}

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Same values as the device-side init (cdae_b200/csrc): element idx of block `which`
 * = float( (2u-1) * 4*sqrt(6/(I+K)) ), u = (word+0.5)*2^-32, word = philox({idx lo, idx hi,
 * which, 0xC0DE}).x — the reference's scale (cdae.hpp:112-113), not its rand() stream. */
void orc_init_params(orc_model* m, uint64_t seed) {
  const double scale = 4. * sqrt(6. / (double)(m->I + m->K));
  const int blocks[3] = {ORC_W, ORC_V, ORC_WU};
  for (int bi = 0; bi < 3; ++bi) {
    int w = blocks[bi];
    if (w == ORC_V && !m->cfg.asymmetric) continue;
    if (w == ORC_WU && !m->cfg.user_factor) continue;
    int64_t n = param_rows(m, w) * param_cols(m, w);
    for (int64_t i = 0; i < n; ++i) {
      uint32_t r[4];
      orc_philox4x32(seed, (uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)w, 0xC0DEu, r);
      double u = ((double)r[0] + 0.5) * (1.0 / 4294967296.0);
      m->p[w][i] = (double)(float)((2. * u - 1.) * scale);
    }
  }
}

/* ------------------------------------------------------------------ losses (loss.hpp) */
double orc_loss_gradient(int lt, double pred, double truth) {
  switch (lt) {
    case ORC_SQUARE: /* loss.hpp:53-55 */
      return -2. * (truth - pred);
    case ORC_LOGISTIC: /* loss.hpp:95-99 (CHECK(pred in (0,1)) aborts in the reference) */
      if (!(pred > 0. && pred < 1.)) return NAN;
      return (pred - truth) / (pred * (1. - pred));
    case ORC_CROSS_ENTROPY: /* loss.hpp:141-147 */
      if (pred < -18) return exp(pred) - truth;
      if (pred > 18) return 1 - truth;
      return 1. / (1. + exp(-pred)) - truth;
    case ORC_LOG: { /* loss.hpp:189-198 */
      double z = pred * truth;
      if (z > 18) return -truth * exp(-z);
      if (z < -18) return -truth;
      return -truth / (1. + exp(z));
    }
    case ORC_LOGM: { /* loss.hpp:239-246 */
      double z = pred;
      if (z > 18) return -truth * exp(-z);
      if (z < -18) return -truth;
      return -truth / (1. + exp(z));
    }
    case ORC_HINGE: { /* loss.hpp:286-291 */
      double z = pred * truth;
      if (z > 1) return 0;
      return -truth;
    }
    case ORC_SQUARED_HINGE: { /* loss.hpp:330-335 */
      double z = pred * truth;
      if (z > 1) return 0;
      return -truth * (1 - z);
    }
  }
  return -2. * (truth - pred); /* Loss::create default: SquareLoss, loss.hpp:364 */
}

double orc_loss_evaluate(int lt, double pred, double truth) {
  switch (lt) {
    case ORC_SQUARE: { /* loss.hpp:48-51 */
      double err = truth - pred;
      return err * err;
    }
    case ORC_LOGISTIC: /* loss.hpp:84-93 */
      if (!(pred >= 0. && pred <= 1.)) return NAN;
      if (truth == 0.) return -log(fmax(0.0001, 1. - pred));
      if (truth == 1.) return -log(fmax(0.0001, pred));
      return 0.;
    case ORC_CROSS_ENTROPY: { /* loss.hpp:132-139 */
      double ret = (1 - truth) * pred;
      if (pred > 18) return ret + exp(-pred);
      if (pred < -18) return ret - pred;
      return ret + log1p(exp(-pred));
    }
    case ORC_LOG: { /* loss.hpp:180-187 */
      double z = pred * truth;
      if (z > 18) return exp(-z);
      if (z < -18) return -z;
      return log1p(exp(-z));
    }
    case ORC_LOGM: { /* loss.hpp:230-237 */
      double z = pred;
      if (z > 18) return truth * exp(-z);
      if (z < -18) return -z * truth;
      return truth * log1p(exp(-pred));
    }
    case ORC_HINGE: { /* loss.hpp:279-284 */
      double z = pred * truth;
      if (z > 1) return 0;
      return 1 - z;
    }
    case ORC_SQUARED_HINGE: { /* loss.hpp:322-328 */
      double z = pred * truth;
      if (z > 1) return 0;
      double d = 1 - z;
      return 0.5 * d * d;
    }
  }
  {
    double err = truth - pred;
    return err * err;
  }
}

/* ------------------------------------------------------------------ forward */
static double act_sigmoid(double x) { /* cdae.hpp:393-401 */
  if (x > 18.) return 1.;
  if (x < -18.) return 0.;
  return 1. / (1. + exp(-x));
}
static double act_tanh(double x) { /* cdae.hpp:403-412 */
  if (x > 9.) return 1.;
  if (x < -9.) return -1.;
  double r = exp(-2. * x);
  return (1. - r) / (1. + r);
}

/* cdae.hpp:373-416 */
void orc_hidden(const orc_model* m, int64_t uid, const int64_t* items, int64_t n, double scale,
                double* z) {
  const int K = m->K;
  const double* W = m->p[ORC_W];
  for (int k = 0; k < K; ++k) z[k] = 0.;
  for (int64_t t = 0; t < n; ++t) { /* :377-380  h1 += W.row(iid) * scale */
    const double* w = W + items[t] * K;
    for (int k = 0; k < K; ++k) z[k] += w[k] * scale;
  }
  if (m->cfg.linear_function) { /* :382-384 */
    const double* uu = m->p[ORC_UU] + uid * K;
    for (int k = 0; k < K; ++k) z[k] = uu[k] * z[k];
  }
  for (int k = 0; k < K; ++k) z[k] += m->p[ORC_B][k]; /* :386 */
  if (m->cfg.user_factor) {                            /* :387-389 */
    const double* wu = m->p[ORC_WU] + uid * K;
    for (int k = 0; k < K; ++k) z[k] += wu[k];
  }
  if (!m->cfg.linear) { /* :391-414 */
    if (!m->cfg.tanh_act)
      for (int k = 0; k < K; ++k) z[k] = act_sigmoid(z[k]);
    else
      for (int k = 0; k < K; ++k) z[k] = act_tanh(z[k]);
  }
}

/* cdae.hpp:418-426 */
double orc_output(const orc_model* m, const double* z, int64_t item) {
  const int K = m->K;
  const double* w = (m->cfg.asymmetric ? m->p[ORC_V] : m->p[ORC_W]) + item * K;
  double dot = 0.;
  for (int k = 0; k < K; ++k) dot += w[k] * z[k];
  double h2 = 0;
  h2 += dot + m->p[ORC_BPRIME][item];
  return h2;
}

/* act'(z) as a function of z, cdae.hpp:208-215 */
static void act_deriv(const orc_model* m, const double* z, double* d) {
  const int K = m->K;
  for (int k = 0; k < K; ++k) d[k] = 1.;
  if (!m->cfg.linear) {
    if (!m->cfg.tanh_act)
      for (int k = 0; k < K; ++k) d[k] = z[k] - z[k] * z[k];
    else
      for (int k = 0; k < K; ++k) d[k] = 1. - z[k] * z[k];
  }
}

/* upd(x, g, a): the AdaGrad / SGD step inlined at every update site, e.g. cdae.hpp:253-257 */
static void upd_row(const orc_model* m, double* x, double* ag, double* grad, int n) {
  if (m->cfg.using_adagrad) {
    for (int k = 0; k < n; ++k) ag[k] += grad[k] * grad[k];
    for (int k = 0; k < n; ++k) grad[k] = grad[k] / (sqrt(ag[k]) + m->cfg.beta);
  }
  for (int k = 0; k < n; ++k) x[k] -= m->cfg.learn_rate * grad[k];
}
static void upd_scalar(const orc_model* m, double* x, double* ag, double grad) { /* :231-236 */
  if (m->cfg.using_adagrad) {
    *ag += grad * grad;
    grad /= (m->cfg.beta + sqrt(*ag));
  }
  *x -= m->cfg.learn_rate * grad;
}

static int in_sorted_or_linear(const int64_t* a, int64_t n, int64_t v) {
  for (int64_t i = 0; i < n; ++i)
    if (a[i] == v) return 1;
  return 0;
}

/* cdae.hpp:198-358 */
void orc_step_sequential(orc_model* m, int64_t uid, const int64_t* in_items, int64_t n_in,
                         const int64_t* negs, int64_t n_negs, const int64_t* out_order) {
  const int K = m->K;
  const orc_config* c = &m->cfg;
  double scale = 1.;
  if (c->scaled) scale /= (1. - c->corruption_ratio); /* :202-205 */

  double* z = (double*)malloc(sizeof(double) * (size_t)K * 6);
  double* d = z + K;    /* z_1_z */
  double* hg = d + K;   /* hidden_gradient */
  double* grad = hg + K;
  double* hd = grad + K; /* hidden_gradient.cwiseProduct(z_1_z) */
  double* uug = hd + K;  /* Uu_grad */
  orc_hidden(m, uid, in_items, n_in, scale, z); /* :207 */
  act_deriv(m, z, d);                           /* :208-215 */
  for (int k = 0; k < K; ++k) hg[k] = 0.;

  const int64_t n_out = m->row_ptr[uid + 1] - m->row_ptr[uid];
  const int32_t* out_csr = m->col + m->row_ptr[uid];
  double* Wd = c->asymmetric ? m->p[ORC_V] : m->p[ORC_W];
  double* Wd_ag = c->asymmetric ? m->p[ORC_V_AG] : m->p[ORC_W_AG];
  double* bp = m->p[ORC_BPRIME];
  double* bp_ag = m->p[ORC_BPRIME_AG];

  /* input_gradient map (:222): dense per-input-slot storage */
  double* ig = (double*)calloc((size_t)(n_in > 0 ? n_in : 1) * (size_t)K, sizeof(double));
  char* ig_set = (char*)calloc((size_t)(n_in > 0 ? n_in : 1), 1);

  for (int64_t t = 0; t < n_out + n_negs; ++t) {
    /* positives loop :225-260 then negatives loop :262-293 */
    const int is_pos = t < n_out;
    const int64_t iid = is_pos ? (out_order ? out_order[t] : (int64_t)out_csr[t]) : negs[t - n_out];
    double y = orc_output(m, z, iid);                               /* :227 / :263 */
    double gradient = orc_loss_gradient(c->loss_type, y, is_pos ? 1. : 0.); /* :228 / :265 */
    upd_scalar(m, &bp[iid], &bp_ag[iid], gradient + c->lambda * bp[iid]);    /* :230-237 */
    double* row = Wd + iid * K;
    for (int k = 0; k < K; ++k) hg[k] += gradient * row[k]; /* :240 / :248 (row before update) */
    int deferred = 0;
    if (is_pos && !c->asymmetric) { /* :249-250 tied & item in corrupted input: defer */
      for (int64_t j = 0; j < n_in; ++j)
        if (in_items[j] == iid) {
          for (int k = 0; k < K; ++k) ig[j * K + k] = gradient * z[k];
          ig_set[j] = 1;
          deferred = 1;
          break;
        }
    }
    if (!deferred) { /* :241-246 / :252-257 / :278-291 */
      for (int k = 0; k < K; ++k) grad[k] = gradient * z[k] + c->lambda * row[k];
      upd_row(m, row, Wd_ag + iid * K, grad, K);
    }
  }

  for (int k = 0; k < K; ++k) hd[k] = hg[k] * d[k];
  if (c->linear_function) { /* :295-299 */
    const double* uu = m->p[ORC_UU] + uid * K;
    for (int k = 0; k < K; ++k) uug[k] = 0. + uu[k] * c->lambda;
  }
  { /* b :301-315 */
    double* b = m->p[ORC_B];
    for (int k = 0; k < K; ++k) grad[k] = hd[k] + c->lambda * b[k];
    upd_row(m, b, m->p[ORC_B_AG], grad, K);
  }
  if (c->user_factor) { /* :317-331 */
    double* wu = m->p[ORC_WU] + uid * K;
    for (int k = 0; k < K; ++k) grad[k] = hd[k] + c->lambda * wu[k];
    upd_row(m, wu, m->p[ORC_WU_AG] + uid * K, grad, K);
  }
  for (int64_t j = 0; j < n_in; ++j) { /* :333-349 */
    double* row = m->p[ORC_W] + in_items[j] * K;
    if (!c->linear_function) {
      for (int k = 0; k < K; ++k) grad[k] = hd[k] * scale + c->lambda * row[k];
    } else {
      const double* uu = m->p[ORC_UU] + uid * K;
      for (int k = 0; k < K; ++k) grad[k] = (uu[k] * hd[k]) * scale + c->lambda * row[k];
      for (int k = 0; k < K; ++k) uug[k] += hd[k] * row[k]; /* :340 */
    }
    if (ig_set[j])
      for (int k = 0; k < K; ++k) grad[k] += ig[j * K + k]; /* :342-343 */
    upd_row(m, row, m->p[ORC_W_AG] + in_items[j] * K, grad, K);
  }
  if (c->linear_function) { /* :351-357 */
    upd_row(m, m->p[ORC_UU] + uid * K, m->p[ORC_UU_AG] + uid * K, uug, K);
  }
  free(ig);
  free(ig_set);
  free(z);
}

/* ------------------------------------------------------------------ frozen batch */
typedef struct {
  double *gW, *gV, *gbp, *gb; /* dense accumulators, zero outside touched rows */
  int64_t *touchW, *touchV, *touchbp;
  int64_t nW, nV, nbp, capW, capV, capbp;
  char *flagW, *flagV, *flagbp;
} frozen_ws;

static void touch(int64_t** list, int64_t* n, int64_t* cap, char* flag, int64_t r) {
  if (flag[r]) return;
  flag[r] = 1;
  if (*n == *cap) {
    *cap = *cap ? *cap * 2 : 1024;
    *list = (int64_t*)realloc(*list, sizeof(int64_t) * (size_t)*cap);
  }
  (*list)[(*n)++] = r;
}

/* ext == NULL: the frozen-batch step.  ext != NULL (data-parallel shard): the item-side
 * gradients are ADDED to ext = [gW (I*K) | gV (I*K, asymmetric only) | gb' (I) | gb (K)] instead
 * of being applied; only the user-private rows (Wu, Uu) are updated here. */
/* bf16 round-to-nearest-even of a value that first goes through fp32 — the operand format of the
 * tcgen05 full-item decode (cdae_b200/csrc/fulldec_tc.cuh). */
static double round_bf16(double x) {
  float f = (float)x;
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (double)f; /* inf / nan */
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&f, &u, 4);
  return (double)f;
}
static int row_contains(const orc_model* m, int64_t uid, int64_t item);

/* full != 0 (H12, no reference function — SURVEY.md F4): the output set of every user is ALL I
 * items, target 1 on the user's train row and 0 elsewhere, i.e. the same step with "negatives =
 * every non-positive once"; negs / neg_ptr are ignored.  rounding != 0 additionally restates the
 * operand rounding of the tensor-core path: z, W' rows and g are rounded to bf16 where they enter
 * the three contractions (y = z.W', hg = sum g W', gW' = sum g z), b' enters y as its bf16 hi+lo
 * pair; products, sums, lambda terms and everything outside the decode stay exact. */
static void step_frozen_impl(orc_model* m, int64_t n_users, const int64_t* uids, const int64_t* in_ptr,
                             const int64_t* in_items, const int64_t* neg_ptr, const int64_t* negs,
                             double* loss_sum_out, double* ext, int full, int rounding) {
  const int K = m->K;
  const orc_config* c = &m->cfg;
  const int64_t I = m->I;
  double scale = 1.;
  if (c->scaled) scale /= (1. - c->corruption_ratio);
  frozen_ws ws;
  memset(&ws, 0, sizeof(ws));
  ws.gW = (double*)calloc((size_t)I * (size_t)K, sizeof(double));
  ws.gV = c->asymmetric ? (double*)calloc((size_t)I * (size_t)K, sizeof(double)) : NULL;
  ws.gbp = (double*)calloc((size_t)I, sizeof(double));
  ws.gb = (double*)calloc((size_t)K, sizeof(double));
  ws.flagW = (char*)calloc((size_t)I, 1);
  ws.flagV = (char*)calloc((size_t)I, 1);
  ws.flagbp = (char*)calloc((size_t)I, 1);
  double* gWu = (double*)calloc((size_t)n_users * (size_t)K, sizeof(double));
  double* gUu = (double*)calloc((size_t)n_users * (size_t)K, sizeof(double));
  double* z = (double*)malloc(sizeof(double) * (size_t)K * 6);
  double* d = z + K;
  double* hg = d + K;
  double* hd = hg + K;
  double* zr = hd + K; /* bf16-rounded z (rounding mode) */
  double* wr = zr + K; /* bf16-rounded W' row */
  double loss_sum = 0.;

  const double* Wd = c->asymmetric ? m->p[ORC_V] : m->p[ORC_W];
  double* gWd = c->asymmetric ? ws.gV : ws.gW;

  for (int64_t ui = 0; ui < n_users; ++ui) {
    const int64_t uid = uids[ui];
    const int64_t* in = in_items + in_ptr[ui];
    const int64_t n_in = in_ptr[ui + 1] - in_ptr[ui];
    const int64_t* ng = full ? NULL : negs + neg_ptr[ui];
    const int64_t n_ng = full ? 0 : neg_ptr[ui + 1] - neg_ptr[ui];
    const int64_t n_out = m->row_ptr[uid + 1] - m->row_ptr[uid];
    const int32_t* out = m->col + m->row_ptr[uid];
    orc_hidden(m, uid, in, n_in, scale, z);
    act_deriv(m, z, d);
    for (int k = 0; k < K; ++k) hg[k] = 0.;
    if (rounding)
      for (int k = 0; k < K; ++k) zr[k] = round_bf16(z[k]);
    const int64_t n_total = full ? I : n_out + n_ng;
    for (int64_t t = 0; t < n_total; ++t) {
      const int is_pos = full ? row_contains(m, uid, t) : t < n_out;
      const int64_t iid = full ? t : (is_pos ? (int64_t)out[t] : ng[t - n_out]);
      const double truth = is_pos ? 1. : 0.;
      const double* row = Wd + iid * K;
      double y;
      if (rounding) {
        const double bp = m->p[ORC_BPRIME][iid], bhi = round_bf16(bp);
        y = 0.;
        for (int k = 0; k < K; ++k) { wr[k] = round_bf16(row[k]); y += zr[k] * wr[k]; }
        y += bhi + round_bf16(bp - bhi);
      } else {
        y = orc_output(m, z, iid);
      }
      double gradient = orc_loss_gradient(c->loss_type, y, truth);
      loss_sum += orc_loss_evaluate(c->loss_type, y, truth);
      if (rounding) gradient = round_bf16(gradient);
      ws.gbp[iid] += gradient + c->lambda * m->p[ORC_BPRIME][iid];
      touch(&ws.touchbp, &ws.nbp, &ws.capbp, ws.flagbp, iid);
      const double* zc = rounding ? zr : z;   /* what enters gW' = sum g z */
      const double* rc = rounding ? wr : row; /* what enters hg = sum g W' */
      for (int k = 0; k < K; ++k) hg[k] += gradient * rc[k];
      /* tied & positive & in the corrupted input: the occurrence merges into the input-row
       * update (one lambda term there, cdae.hpp:249-250,342-343); else lambda term here. */
      int merged = is_pos && !c->asymmetric && in_sorted_or_linear(in, n_in, iid);
      double* g = gWd + iid * K;
      if (merged)
        for (int k = 0; k < K; ++k) g[k] += gradient * zc[k];
      else
        for (int k = 0; k < K; ++k) g[k] += gradient * zc[k] + c->lambda * row[k];
      if (c->asymmetric)
        touch(&ws.touchV, &ws.nV, &ws.capV, ws.flagV, iid);
      else
        touch(&ws.touchW, &ws.nW, &ws.capW, ws.flagW, iid);
    }
    for (int k = 0; k < K; ++k) hd[k] = hg[k] * d[k];
    for (int k = 0; k < K; ++k) ws.gb[k] += hd[k] + c->lambda * m->p[ORC_B][k];
    if (c->user_factor) {
      const double* wu = m->p[ORC_WU] + uid * K;
      for (int k = 0; k < K; ++k) gWu[ui * K + k] = hd[k] + c->lambda * wu[k];
    }
    if (c->linear_function) {
      const double* uu = m->p[ORC_UU] + uid * K;
      for (int k = 0; k < K; ++k) gUu[ui * K + k] = 0. + uu[k] * c->lambda;
    }
    for (int64_t j = 0; j < n_in; ++j) {
      const double* row = m->p[ORC_W] + in[j] * K;
      double* g = ws.gW + in[j] * K;
      if (!c->linear_function) {
        for (int k = 0; k < K; ++k) g[k] += hd[k] * scale + c->lambda * row[k];
      } else {
        const double* uu = m->p[ORC_UU] + uid * K;
        for (int k = 0; k < K; ++k) g[k] += (uu[k] * hd[k]) * scale + c->lambda * row[k];
        for (int k = 0; k < K; ++k) gUu[ui * K + k] += hd[k] * row[k];
      }
      touch(&ws.touchW, &ws.nW, &ws.capW, ws.flagW, in[j]);
    }
  }
  if (ext) {
    double* e = ext;
    for (int64_t i = 0; i < I * K; ++i) e[i] += ws.gW[i];
    e += I * K;
    if (c->asymmetric) {
      for (int64_t i = 0; i < I * K; ++i) e[i] += ws.gV[i];
      e += I * K;
    }
    for (int64_t i = 0; i < I; ++i) e[i] += ws.gbp[i];
    e += I;
    for (int k = 0; k < K; ++k) e[k] += ws.gb[k];
    ws.nW = ws.nV = ws.nbp = 0; /* nothing item-side is applied below */
  }
  /* one upd per touched row */
  for (int64_t t = 0; t < ws.nW; ++t) {
    int64_t r = ws.touchW[t];
    upd_row(m, m->p[ORC_W] + r * K, m->p[ORC_W_AG] + r * K, ws.gW + r * K, K);
  }
  for (int64_t t = 0; t < ws.nV; ++t) {
    int64_t r = ws.touchV[t];
    upd_row(m, m->p[ORC_V] + r * K, m->p[ORC_V_AG] + r * K, ws.gV + r * K, K);
  }
  for (int64_t t = 0; t < ws.nbp; ++t) {
    int64_t r = ws.touchbp[t];
    upd_scalar(m, &m->p[ORC_BPRIME][r], &m->p[ORC_BPRIME_AG][r], ws.gbp[r]);
  }
  if (n_users > 0 && !ext) upd_row(m, m->p[ORC_B], m->p[ORC_B_AG], ws.gb, K);
  for (int64_t ui = 0; ui < n_users; ++ui) {
    if (c->user_factor)
      upd_row(m, m->p[ORC_WU] + uids[ui] * K, m->p[ORC_WU_AG] + uids[ui] * K, gWu + ui * K, K);
    if (c->linear_function)
      upd_row(m, m->p[ORC_UU] + uids[ui] * K, m->p[ORC_UU_AG] + uids[ui] * K, gUu + ui * K, K);
  }
  if (loss_sum_out) *loss_sum_out = loss_sum;
  free(ws.gW); free(ws.gV); free(ws.gbp); free(ws.gb);
  free(ws.flagW); free(ws.flagV); free(ws.flagbp);
  free(ws.touchW); free(ws.touchV); free(ws.touchbp);
  free(gWu); free(gUu); free(z);
}

void orc_step_frozen(orc_model* m, int64_t n_users, const int64_t* uids, const int64_t* in_ptr,
                     const int64_t* in_items, const int64_t* neg_ptr, const int64_t* negs,
                     double* loss_sum_out) {
  step_frozen_impl(m, n_users, uids, in_ptr, in_items, neg_ptr, negs, loss_sum_out, NULL, 0, 0);
}

void orc_step_frozen_full(orc_model* m, int64_t n_users, const int64_t* uids, const int64_t* in_ptr,
                          const int64_t* in_items, int rounding, double* loss_sum_out) {
  step_frozen_impl(m, n_users, uids, in_ptr, in_items, NULL, NULL, loss_sum_out, NULL, 1, rounding);
}
void orc_shard_gradients_full(orc_model* m, int64_t n_users, const int64_t* uids, const int64_t* in_ptr,
                              const int64_t* in_items, int rounding, double* loss_sum_out,
                              double* dense_grad) {
  step_frozen_impl(m, n_users, uids, in_ptr, in_items, NULL, NULL, loss_sum_out, dense_grad, 1, rounding);
}

void orc_shard_gradients(orc_model* m, int64_t n_users, const int64_t* uids, const int64_t* in_ptr,
                         const int64_t* in_items, const int64_t* neg_ptr, const int64_t* negs,
                         double* loss_sum_out, double* dense_grad) {
  step_frozen_impl(m, n_users, uids, in_ptr, in_items, neg_ptr, negs, loss_sum_out, dense_grad, 0, 0);
}

/* One upd per element whose summed gradient is non-zero (upd with g = 0 is a no-op, so this
 * equals "one upd per touched row"); any_steps = 0 skips b (no user contributed). */
void orc_apply_dense(orc_model* m, const double* dense_grad, int any_steps) {
  const int K = m->K;
  const int64_t I = m->I;
  const double* e = dense_grad;
  double* g = (double*)malloc(sizeof(double) * (size_t)K);
  for (int64_t r = 0; r < I; ++r) {
    int nz = 0;
    for (int k = 0; k < K; ++k) { g[k] = e[r * K + k]; nz |= g[k] != 0.; }
    if (nz) upd_row(m, m->p[ORC_W] + r * K, m->p[ORC_W_AG] + r * K, g, K);
  }
  e += I * K;
  if (m->cfg.asymmetric) {
    for (int64_t r = 0; r < I; ++r) {
      int nz = 0;
      for (int k = 0; k < K; ++k) { g[k] = e[r * K + k]; nz |= g[k] != 0.; }
      if (nz) upd_row(m, m->p[ORC_V] + r * K, m->p[ORC_V_AG] + r * K, g, K);
    }
    e += I * K;
  }
  for (int64_t r = 0; r < I; ++r)
    if (e[r] != 0.) upd_scalar(m, &m->p[ORC_BPRIME][r], &m->p[ORC_BPRIME_AG][r], e[r]);
  e += I;
  if (any_steps) {
    for (int k = 0; k < K; ++k) g[k] = e[k];
    upd_row(m, m->p[ORC_B], m->p[ORC_B_AG], g, K);
  }
  free(g);
}

/* ------------------------------------------------------------------ recommend */
static int row_contains(const orc_model* m, int64_t uid, int64_t item) {
  const int32_t* a = m->col + m->row_ptr[uid];
  int64_t n = m->row_ptr[uid + 1] - m->row_ptr[uid];
  if (m->rows_sorted) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (a[mid] < item) lo = mid + 1; else hi = mid;
    }
    return lo < n && a[lo] == item;
  }
  for (int64_t i = 0; i < n; ++i)
    if (a[i] == item) return 1;
  return 0;
}

/* cdae.hpp:162-196 + heap.hpp:44-52 (+ utils.hpp:15-19): bounded min-heap on score, a new
 * candidate replaces the current minimum only on STRICT improvement, items visited in
 * ascending id, result sorted by score descending. */
int orc_recommend(const orc_model* m, int64_t uid, int64_t topk, int64_t* ids_out,
                  double* scores_out) {
  const int K = m->K;
  double* z = (double*)malloc(sizeof(double) * (size_t)K);
  const int64_t n_u = m->row_ptr[uid + 1] - m->row_ptr[uid];
  if (m->cfg.corruption_ratio != 1.) { /* :168-172 (uncorrupted, scale = 1) */
    int64_t* items = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_u > 0 ? n_u : 1));
    for (int64_t i = 0; i < n_u; ++i) items[i] = m->col[m->row_ptr[uid] + i];
    orc_hidden(m, uid, items, n_u, 1.0, z);
    free(items);
  } else {
    orc_hidden(m, uid, NULL, 0, 1.0, z);
  }
  char* rated = (char*)calloc((size_t)m->I, 1);
  for (int64_t i = 0; i < n_u; ++i) rated[m->col[m->row_ptr[uid] + i]] = 1;
  int64_t size = 0;
  for (int64_t item = 0; item < m->I; ++item) {
    if (rated[item]) continue; /* :177-179 */
    double pred = orc_output(m, z, item);
    if (size < topk) {
      ids_out[size] = item;
      scores_out[size] = pred;
      ++size;
    } else {
      /* current minimum; among equal minima evict the highest id (the reference's choice
       * among exact ties is whatever std::pop_heap surfaces — unspecified) */
      int64_t mn = 0;
      for (int64_t t = 1; t < size; ++t)
        if (scores_out[t] < scores_out[mn] ||
            (scores_out[t] == scores_out[mn] && ids_out[t] > ids_out[mn]))
          mn = t;
      if (pred > scores_out[mn]) { /* heap.hpp:45 comp_(t, front): strict */
        ids_out[mn] = item;
        scores_out[mn] = pred;
      }
    }
  }
  free(rated);
  free(z);
  if (size != topk) return -1; /* CHECK_EQ(topk_heap.size(), topk) :187 */
  /* sort by score desc, id asc on ties (insertion sort, k is ~10) */
  for (int64_t a = 1; a < size; ++a) {
    int64_t id = ids_out[a];
    double s = scores_out[a];
    int64_t b = a - 1;
    while (b >= 0 && (scores_out[b] < s || (scores_out[b] == s && ids_out[b] > id))) {
      ids_out[b + 1] = ids_out[b];
      scores_out[b + 1] = scores_out[b];
      --b;
    }
    ids_out[b + 1] = id;
    scores_out[b + 1] = s;
  }
  return 0;
}

/* cdae.hpp:78-101, one corruption, explicit keep mask (NULL = keep all) */
double orc_data_loss(const orc_model* m, const uint8_t* keep) {
  const int K = m->K;
  double scale = 1;
  if (m->cfg.scaled) scale /= (1. - m->cfg.corruption_ratio);
  double* z = (double*)malloc(sizeof(double) * (size_t)K);
  double rets = 0.;
  int64_t maxn = 1;
  for (int64_t u = 0; u < m->U; ++u)
    if (m->row_ptr[u + 1] - m->row_ptr[u] > maxn) maxn = m->row_ptr[u + 1] - m->row_ptr[u];
  int64_t* in = (int64_t*)malloc(sizeof(int64_t) * (size_t)maxn);
  for (int64_t u = 0; u < m->U; ++u) {
    int64_t s0 = m->row_ptr[u], s1 = m->row_ptr[u + 1], n_in = 0;
    for (int64_t s = s0; s < s1; ++s)
      if (!keep || keep[s]) in[n_in++] = m->col[s];
    orc_hidden(m, u, in, n_in, scale, z);
    double user_rets = 0;
    for (int64_t s = s0; s < s1; ++s)
      user_rets += orc_loss_evaluate(m->cfg.loss_type, orc_output(m, z, m->col[s]), 1.);
    rets = rets + user_rets / 1.0;
  }
  free(in);
  free(z);
  return rets;
}

/* cdae.hpp:103-107 with penalty.hpp:36-39 (L2: squaredNorm; empty matrix -> 0) */
double orc_penalty_loss(const orc_model* m) {
  const int blocks[5] = {ORC_W, ORC_V, ORC_WU, ORC_B, ORC_BPRIME};
  double tot = 0.;
  for (int bi = 0; bi < 5; ++bi) {
    int w = blocks[bi];
    if (w == ORC_V && !m->cfg.asymmetric) continue; /* V stays 0x0 */
    if (w == ORC_WU && !m->cfg.user_factor) continue;
    int64_t n = param_rows(m, w) * param_cols(m, w);
    double s = 0.;
    for (int64_t i = 0; i < n; ++i) s += m->p[w][i] * m->p[w][i];
    tot += s;
  }
  return 0.5 * m->cfg.lambda * tot;
}

/* cdae.hpp:148-159 */
void orc_user_representations(const orc_model* m, double* out) {
  int64_t maxn = 1;
  for (int64_t u = 0; u < m->U; ++u)
    if (m->row_ptr[u + 1] - m->row_ptr[u] > maxn) maxn = m->row_ptr[u + 1] - m->row_ptr[u];
  int64_t* in = (int64_t*)malloc(sizeof(int64_t) * (size_t)maxn);
  for (int64_t u = 0; u < m->U; ++u) {
    int64_t n = m->row_ptr[u + 1] - m->row_ptr[u];
    for (int64_t i = 0; i < n; ++i) in[i] = m->col[m->row_ptr[u] + i];
    orc_hidden(m, u, in, n, 1.0, out + u * m->K);
  }
  free(in);
}

/* evaluation.hpp:183-219 */
void orc_evaluate_rec_list(const int64_t* list, int64_t n_list, const int64_t* test_items,
                           int64_t n_test, double* rets) {
  for (int i = 0; i < 8; ++i) rets[i] = 0.;
  int64_t TOPK = 20;
  double hit = 0., map5 = 0, map10 = 0;
  if (n_list < TOPK) TOPK = n_list;
  for (int64_t idx = 0; idx < TOPK; ++idx) {
    if (in_sorted_or_linear(test_items, n_test, list[idx])) {
      hit += 1.;
      if (idx < 5) map5 += hit / (double)(idx + 1);
      if (idx < 10) map10 += hit / (double)(idx + 1);
    }
    if (idx == 0) {
      rets[0] = hit / 1.;
      rets[3] = hit / (double)n_test;
    } else if (idx == 4) {
      rets[1] = hit / 5.;
      rets[4] = hit / (double)n_test;
    } else if (idx == 9) {
      rets[2] = hit / 10.;
      rets[5] = hit / (double)n_test;
    }
  }
  rets[6] = map5 / (double)(n_test < 5 ? n_test : 5);
  rets[7] = map10 / (double)(n_test < 10 ? n_test : 10);
}

/* evaluation.hpp:113-181 (list length fixed at 10, :145; mean over users that have test items) */
int64_t orc_topn_evaluate(const orc_model* m, const int64_t* trp, const int32_t* tcol,
                          double* out8) {
  int64_t n_test_users = 0;
  for (int64_t u = 0; u < m->U; ++u)
    if (trp[u + 1] > trp[u]) ++n_test_users;
  for (int i = 0; i < 8; ++i) out8[i] = 0.;
  for (int64_t u = 0; u < m->U; ++u) {
    int64_t nt = trp[u + 1] - trp[u];
    if (nt == 0) continue;
    int64_t ids[10];
    double sc[10], r8[8];
    int64_t* t = (int64_t*)malloc(sizeof(int64_t) * (size_t)nt);
    for (int64_t i = 0; i < nt; ++i) t[i] = tcol[trp[u] + i];
    if (orc_recommend(m, u, 10, ids, sc) != 0) {
      free(t);
      return -1;
    }
    orc_evaluate_rec_list(ids, 10, t, nt, r8);
    for (int i = 0; i < 8; ++i) out8[i] += r8[i] / (double)n_test_users; /* :162-166 */
    free(t);
  }
  return n_test_users;
}

/* ------------------------------------------------------------------ sampling spec */
void orc_sample_keep(const orc_model* m, uint64_t seed, uint32_t pass, int64_t uid,
                     uint8_t* keep) {
  const double q = m->cfg.corruption_ratio;
  const int64_t n = m->row_ptr[uid + 1] - m->row_ptr[uid];
  if (q <= 0.) {
    for (int64_t s = 0; s < n; ++s) keep[s] = 1;
    return;
  }
  if (q >= 1.) {
    for (int64_t s = 0; s < n; ++s) keep[s] = 0;
    return;
  }
  const uint32_t thr = (uint32_t)floor(q * 4294967296.0);
  for (int64_t s = 0; s < n; ++s) {
    uint32_t r[4];
    orc_philox4x32(seed, (uint32_t)uid, (uint32_t)(s >> 2), pass, 0u, r);
    keep[s] = r[s & 3] > thr; /* cdae.hpp:366: keep iff uniform > ratio */
  }
}

/* recsys_model_base.hpp:46-57: uniform over items, redraw while it is one of the user's
 * positives; with replacement; n_u * num_neg draws (cdae.hpp:217-220) */
void orc_sample_negatives(const orc_model* m, uint64_t seed, uint32_t pass, int64_t uid,
                          int64_t* negs) {
  const int64_t n = (m->row_ptr[uid + 1] - m->row_ptr[uid]) * m->cfg.num_neg;
  for (int64_t d = 0; d < n; ++d) {
    uint32_t r[4];
    orc_philox4x32(seed, (uint32_t)uid, (uint32_t)(d >> 2), pass, 1u, r);
    int64_t item = (int64_t)(((uint64_t)r[d & 3] * (uint64_t)m->I) >> 32);
    for (uint32_t a = 1; row_contains(m, uid, item); ++a) {
      orc_philox4x32(seed, (uint32_t)uid, (uint32_t)d, pass, 2u + ((a - 1) >> 2), r);
      item = (int64_t)(((uint64_t)r[(a - 1) & 3] * (uint64_t)m->I) >> 32);
    }
    negs[d] = item;
  }
}

/* cdae.hpp:136-146 */
double orc_train_epoch(orc_model* m, uint64_t seed, int64_t epoch, int64_t batch_users, int64_t u0,
                       int64_t u1) {
  const int cnum = m->cfg.num_corruptions;
  const int nu = m->cfg.num_neg;
  double loss_total = 0.;
  if (batch_users < 1) batch_users = 1;
  int64_t maxn = 1;
  for (int64_t u = u0; u < u1; ++u)
    if (m->row_ptr[u + 1] - m->row_ptr[u] > maxn) maxn = m->row_ptr[u + 1] - m->row_ptr[u];
  if (batch_users == 1) {
    uint8_t* keep = (uint8_t*)malloc((size_t)maxn);
    int64_t* in = (int64_t*)malloc(sizeof(int64_t) * (size_t)maxn);
    int64_t* negs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(maxn * (nu > 0 ? nu : 1)));
    for (int64_t u = u0; u < u1; ++u) {
      const int64_t s0 = m->row_ptr[u], n = m->row_ptr[u + 1] - s0;
      for (int cidx = 0; cidx < cnum; ++cidx) {
        uint32_t pass = (uint32_t)(epoch * cnum + cidx);
        orc_sample_keep(m, seed, pass, u, keep);
        int64_t n_in = 0;
        for (int64_t s = 0; s < n; ++s)
          if (keep[s]) in[n_in++] = m->col[s0 + s];
        orc_sample_negatives(m, seed, pass, u, negs);
        orc_step_sequential(m, u, in, n_in, negs, n * nu, NULL);
      }
    }
    free(keep); free(in); free(negs);
    return 0.;
  }
  for (int64_t b0 = u0; b0 < u1; b0 += batch_users) {
    int64_t b1 = b0 + batch_users < u1 ? b0 + batch_users : u1;
    int64_t nb = b1 - b0;
    int64_t nnz = m->row_ptr[b1] - m->row_ptr[b0];
    int64_t* uids = (int64_t*)malloc(sizeof(int64_t) * (size_t)nb);
    int64_t* in_ptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nb + 1));
    int64_t* neg_ptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nb + 1));
    int64_t* in = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t* negs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz * nu > 0 ? nnz * nu : 1));
    uint8_t* keep = (uint8_t*)malloc((size_t)maxn);
    for (int cidx = 0; cidx < cnum; ++cidx) {
      uint32_t pass = (uint32_t)(epoch * cnum + cidx);
      in_ptr[0] = 0;
      neg_ptr[0] = 0;
      for (int64_t i = 0; i < nb; ++i) {
        int64_t u = b0 + i;
        uids[i] = u;
        const int64_t s0 = m->row_ptr[u], n = m->row_ptr[u + 1] - s0;
        orc_sample_keep(m, seed, pass, u, keep);
        int64_t n_in = 0;
        for (int64_t s = 0; s < n; ++s)
          if (keep[s]) in[in_ptr[i] + n_in++] = m->col[s0 + s];
        in_ptr[i + 1] = in_ptr[i] + n_in;
        orc_sample_negatives(m, seed, pass, u, negs + neg_ptr[i]);
        neg_ptr[i + 1] = neg_ptr[i] + n * nu;
      }
      double ls = 0.;
      orc_step_frozen(m, nb, uids, in_ptr, in, neg_ptr, negs, &ls);
      loss_total += ls;
    }
    free(uids); free(in_ptr); free(neg_ptr); free(in); free(negs); free(keep);
  }
  return loss_total;
}

/* One epoch of full-item-decode training (H12): frozen minibatches of batch_users consecutive
 * users, Philox keep masks as in orc_train_epoch, no negatives. */
double orc_train_epoch_full(orc_model* m, uint64_t seed, int64_t epoch, int64_t batch_users,
                            int64_t u0, int64_t u1, int rounding) {
  const int cnum = m->cfg.num_corruptions;
  double loss_total = 0.;
  if (batch_users < 1) batch_users = 1;
  int64_t maxn = 1;
  for (int64_t u = u0; u < u1; ++u)
    if (m->row_ptr[u + 1] - m->row_ptr[u] > maxn) maxn = m->row_ptr[u + 1] - m->row_ptr[u];
  uint8_t* keep = (uint8_t*)malloc((size_t)maxn);
  for (int64_t b0 = u0; b0 < u1; b0 += batch_users) {
    const int64_t b1 = b0 + batch_users < u1 ? b0 + batch_users : u1, nb = b1 - b0;
    const int64_t nnz = m->row_ptr[b1] - m->row_ptr[b0];
    int64_t* uids = (int64_t*)malloc(sizeof(int64_t) * (size_t)nb);
    int64_t* in_ptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nb + 1));
    int64_t* in = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    for (int cidx = 0; cidx < cnum; ++cidx) {
      const uint32_t pass = (uint32_t)(epoch * cnum + cidx);
      in_ptr[0] = 0;
      for (int64_t i = 0; i < nb; ++i) {
        const int64_t u = b0 + i, s0 = m->row_ptr[u], n = m->row_ptr[u + 1] - s0;
        uids[i] = u;
        orc_sample_keep(m, seed, pass, u, keep);
        int64_t n_in = 0;
        for (int64_t s = 0; s < n; ++s)
          if (keep[s]) in[in_ptr[i] + n_in++] = m->col[s0 + s];
        in_ptr[i + 1] = in_ptr[i] + n_in;
      }
      double ls = 0.;
      orc_step_frozen_full(m, nb, uids, in_ptr, in, rounding, &ls);
      loss_total += ls;
    }
    free(uids); free(in_ptr); free(in);
  }
  free(keep);
  return loss_total;
}

/* OUR multi-core variant (the reference trains on one thread): users are
 * split statically across threads, each runs the sequential step on the shared parameters
 * without locks (Hogwild).  Only used as the "stronger CPU baseline" in bench reports. */
typedef struct {
  orc_model* m;
  uint64_t seed;
  int64_t epoch, a, b;
} hogwild_arg;
static void* hogwild_worker(void* p) {
  hogwild_arg* h = (hogwild_arg*)p;
  orc_train_epoch(h->m, h->seed, h->epoch, 1, h->a, h->b);
  return NULL;
}
double orc_train_epoch_hogwild(orc_model* m, uint64_t seed, int64_t epoch, int64_t u0, int64_t u1,
                               int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  hogwild_arg args[256];
  const int64_t len = u1 - u0;
  for (int t = 0; t < n_threads; ++t) {
    args[t].m = m;
    args[t].seed = seed;
    args[t].epoch = epoch;
    args[t].a = u0 + (len * t) / n_threads;
    args[t].b = u0 + (len * (t + 1)) / n_threads;
    pthread_create(&th[t], NULL, hogwild_worker, &args[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  return 0.;
}
