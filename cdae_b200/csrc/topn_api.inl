// cdae_b200/csrc/topn_api.inl — cdae_topn_* entry points (included at the end of api.cu).

// Uncorrupted hidden vectors of this rank's users into topn_z (cdae.hpp:167-172: scale = 1;
// with corruption_ratio == 1 the reference feeds an EMPTY input set).
static int encode_all_users(cdae_handle* h) {
  if (!h->plan_valid) TRY(build_plan(h));
  TRY(ensure_scratch(h, h->plan_max_users, h->plan_max_slots));
  TRY(ensure(h, h->topn_z, (size_t)(h->U * h->ld)));
  // ONE gather over every input chunk of this rank (the work-item list is contiguous across
  // minibatches) accumulating straight into topn_z rows addressed by global uid, then ONE
  // in-place activate: 100,000 users per launch instead of 8192 keeps the gather near its
  // bandwidth bound.
  int64_t n_in = 0, n_users = 0;
  for (const MiniBatch& p : h->plan) { n_in += p.n_in; n_users += p.n_users; }
  if (n_users == 0) return 0;
  CU(cudaMemsetAsync(h->topn_z.p, 0, sizeof(float) * (size_t)(h->U * h->ld), h->stream));
  BatchDev bt = make_batch(h, h->plan_in.p, n_in, h->plan_out.p, 0, h->plan_uids.p, n_users);
  bt.H = bt.Z = h->topn_z.p;
  bt.flags = BATCH_BY_UID | BATCH_KEEP_ALL;
  if (h->cfg.corruption_ratio != 1.) TRY(launch_gather(h, bt, nullptr, false));  // q == 1: empty input set
  TRY(launch_activate(h, bt, 1.f));
  return 0;
}

// ---- tensor-core candidate phase (topn_tc.cuh) ---------------------------------------------
static int tc_make_map(CUtensorMap* m, void* base, uint64_t rows, uint64_t Kp, uint32_t box_rows) {
  cuuint64_t dims[2] = {Kp, rows};
  cuuint64_t strides[1] = {Kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)tc::KBLK, box_rows};
  cuuint32_t es[2] = {1, 1};
  // the driver entry point is fetched through the runtime so the library has no link-time
  // dependency on libcuda.so.1 (it must load, for symbol checks, on machines without a driver)
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return set_error(CDAE_E_CUDA, "driver lacks cuTensorMapEncodeTiled");
    encode = (encode_fn)fn;
  }
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(CDAE_E_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
  return 0;
}

static bool tc_path_wanted(cdae_handle* h, int topk) {
  const char* e = getenv("CDAE_B200_TOPN");  // "fp32" forces the CUDA-core path, "tc" is the default
  if (e && strcmp(e, "fp32") == 0) return false;
  return h->K + 2 <= tc::MAX_KB * tc::KBLK && topk <= 16;
}

template <int KB>
static int tc_launch(cdae_handle* h, const CUtensorMap& ma, const CUtensorMap& mb, const tc::TcArgs& a, int grid_x) {
  const dim3 grid(grid_x, a.n_splits);
  const size_t dyn = tc::smem_bytes(KB);
  {  // per launch: the attribute belongs to the CURRENT device (several handles / devices per process)
    CU(cudaFuncSetAttribute(tc::topn_tc_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  }
  tc::topn_tc_kernel<KB><<<grid, tc::n_threads(KB), dyn, h->stream>>>(ma, mb, a);
  return 0;
}

// One sweep of the tensor-core candidate kernel over `n_users` users (`users` = their global ids, or
// nullptr for 0..n-1) followed by the verifying re-rank.  Users whose list the bound cannot prove
// are written to redo_list[0 .. *n_redo) together with a start threshold for a second sweep.
// Wb (the packed item side) must already be in h->tc_wb; init_thr is nullable.
static int tc_pass(cdae_handle* h, const float* Wd, const int32_t* users, int64_t n_users, int topk,
                   const float* init_thr, int32_t* redo_list, float* redo_thr, int* redo_cnt, int* n_redo,
                   bool allow_split = true) {
  const int K = h->K, Kp = (int)round_up(K + 2, tc::KBLK), KB = Kp / tc::KBLK;
  const int64_t n_pad = round_up(n_users, tc::TILE_U), I_pad = round_up(h->I, tc::TILE_I);
  TRY(ensure(h, h->tc_zb, (size_t)(n_pad * Kp)));
  CU(cudaMemsetAsync(redo_cnt, 0, sizeof(int), h->stream));
  {
    ProfScope ps(h, CDAE_K_TOPN_PACK);
    tc::pack_z_bf16_kernel<<<cdiv(n_pad * 32, 256), 256, 0, h->stream>>>(
        h->topn_z.p, users, (int)n_users, n_pad, K, h->ld, Kp, h->tc_wmax.p,
        reinterpret_cast<__nv_bfloat16*>(h->tc_zb.p), h->tc_eps.p);
    KERNEL_OK(h);
  }
  alignas(64) CUtensorMap ma, mb;
  TRY(tc_make_map(&ma, h->tc_zb.p, (uint64_t)n_pad, (uint64_t)Kp, tc::TILE_U));
  TRY(tc_make_map(&mb, h->tc_wb.p, (uint64_t)I_pad, (uint64_t)Kp, tc::TILE_I));
  tc::TcArgs a;
  a.n_users = (int)n_users; a.I = h->I; a.n_tiles = (int)(I_pad / tc::TILE_I);
  a.users = users; a.row_ptr = h->row_ptr_d.p; a.col = h->col_d.p;
  a.cand_id = h->cand_id.p; a.cand_s = h->cand_s.p; a.cand_cnt = h->cand_cnt.p; a.cand_thr = h->tc_thr.p;
  a.init_thr = init_thr;
  // Dense rows (more than ~half a rated item per user per 256-item tile, e.g. config C's 145 items over
  // 106 tiles): walking the CSR inside the kernel is one dependent global load per item and paces the
  // whole kernel (profiles/r01_h_*: 139 TFLOP/s at config C against 965 at config D), so the rated sets
  // become a bitmap first.  Sparse rows keep the in-kernel walk (the bitmap would cost more than it saves).
  a.bits = nullptr;
  a.words = I_pad / 32;
  const double per_tile = (double)h->nnz / (double)std::max<int64_t>(h->U, 1) / (double)(I_pad / tc::TILE_I);
  static const char* bm_env = getenv("CDAE_B200_TOPN_BITMAP");      // "0" / "1" force the choice (A/B runs)
  const bool want_bits = bm_env ? atoi(bm_env) != 0 : per_tile > 0.5;
  if (want_bits && (size_t)n_pad * (size_t)a.words * 4 <= ((size_t)8 << 30)) {
    TRY(ensure(h, h->fd_bits, (size_t)(n_pad * a.words)));
    CU(cudaMemsetAsync(h->fd_bits.p, 0, sizeof(uint32_t) * (size_t)(n_pad * a.words), h->stream));
    ProfScope ps(h, CDAE_K_TOPN_PACK);
    tc::topn_bitmap_kernel<<<cdiv(n_users * 32, 256), 256, 0, h->stream>>>(users, (int)n_users, h->row_ptr_d.p, h->col_d.p,
                                                                          h->I, a.words, h->fd_bits.p);
    KERNEL_OK(h);
    a.bits = h->fd_bits.p;
  }
  const int grid = (int)(n_pad / tc::TILE_U);
  // Item ranges: a sweep that starts from known thresholds (init_thr) is paced by the MMA, not by
  // candidate handling, and has few user tiles — cut the items into ranges while the grid still fits one wave.
  int S = 1;
  // (not for a first sweep that starts from probe thresholds: its candidate lists need whole segments)
  if (init_thr && allow_split) while (S < 8 && grid * S * 2 <= h->sm_count && a.n_tiles / (S * 2) >= 8) S *= 2;
  a.n_splits = S;
  a.tiles_per_split = (a.n_tiles + S - 1) / S;
  a.seg = tc::CAND_MAX / S;
  TRY(ensure(h, h->cand_cnt, (size_t)(n_users * S)));
  TRY(ensure(h, h->tc_thr, (size_t)(n_users * S)));
  a.cand_cnt = h->cand_cnt.p; a.cand_thr = h->tc_thr.p;
  {
    ProfScope ps(h, CDAE_K_TOPN);
    switch (KB) {
      case 1: TRY(tc_launch<1>(h, ma, mb, a, grid)); break;
      case 2: TRY(tc_launch<2>(h, ma, mb, a, grid)); break;
      case 3: TRY(tc_launch<3>(h, ma, mb, a, grid)); break;
      case 4: TRY(tc_launch<4>(h, ma, mb, a, grid)); break;
      default: TRY(tc_launch<5>(h, ma, mb, a, grid)); break;
    }
    KERNEL_OK(h);
  }
  {
    ProfScope ps(h, CDAE_K_TOPN_RERANK);
    topn_rerank_kernel<<<cdiv(n_users, 8), 256, 0, h->stream>>>(
        h->topn_z.p, Wd, h->m.bp, h->K, h->ld, users, (int)n_users, h->cand_id.p, h->cand_cnt.p, a.seg, S,
        topk, h->topn_ids.p, h->topn_scores.p, h->flag_d.p, h->tc_thr.p, h->tc_eps.p, redo_list, redo_cnt,
        redo_thr);
    KERNEL_OK(h);
  }
  CU(cudaMemcpyAsync(n_redo, redo_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

// Probe pass (topn_tc.cuh, "Probe pass"): start thresholds for sweep 1 from the M items with the largest
// mean-user score.  Leaves them in h->tc_probe_thr (one per user of the list) and returns M in *m_out, or
// *m_out = 0 when the item table is too small for a probe to pay (fewer than 512 items) or
// CDAE_B200_TOPN_PROBE=0 (1 = on; n >= 2 = on with up to n tiles of 256 probe items instead of 4).  h->tc_wb must hold the packed item side.
#ifndef TOPN_PROBE_DEFAULT
#define TOPN_PROBE_DEFAULT 1   // measured at config B: candidate phase 2.14 -> 1.69 ms, all 100,000 lists verified and equal
#endif                         // to the lists without it (profiles/r02_n_topn_probe_ab.jsonl)
static int tc_probe_thresholds(cdae_handle* h, const float* Wd, const int32_t* users, int64_t n_users, int topk, int* m_out) {
  *m_out = 0;
  const char* env = getenv("CDAE_B200_TOPN_PROBE");   // read per call: tests switch it inside one process
  const int knob = env ? atoi(env) : TOPN_PROBE_DEFAULT;     // 0 off, 1 on (up to 4 tiles of 256 items), n >= 2: up to n tiles
  if (knob <= 0 || h->I < 512 || n_users <= 0 || h->I >= 0x7fffffff) return 0;
  const int K = h->K, Kp = (int)round_up(K + 2, tc::KBLK), KB = Kp / tc::KBLK;
  const int max_tiles = knob == 1 ? 4 : std::min(knob, 32);
  const int M = tc::TILE_I * (int)std::min<int64_t>(max_tiles, std::max<int64_t>(1, h->I / 2048));   // 256 .., <= I / 2
  const int64_t n_pad = round_up(n_users, tc::TILE_U);
  const int64_t words = M / 32;
  const int I = (int)h->I;
  TRY(ensure(h, h->tc_probe_zsum, (size_t)h->ld));
  TRY(ensure(h, h->tc_probe_keys, (size_t)(2 * h->I)));
  TRY(ensure(h, h->tc_probe_ids, (size_t)(2 * h->I)));
  TRY(ensure(h, h->tc_probe_pos, (size_t)h->I));
  TRY(ensure(h, h->tc_probe_wb, (size_t)M * Kp));
  TRY(ensure(h, h->tc_probe_bits, (size_t)(n_pad * words)));
  TRY(ensure(h, h->tc_probe_thr, (size_t)n_users));
  TRY(ensure(h, h->tc_zb, (size_t)(n_pad * Kp)));
  float* keys_in = h->tc_probe_keys.p;
  float* keys_out = h->tc_probe_keys.p + h->I;
  int32_t* ids_in = h->tc_probe_ids.p;
  int32_t* ids_out = h->tc_probe_ids.p + h->I;
  size_t tmp_bytes = 0;
  if (cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, keys_in, keys_out, ids_in, ids_out, I, 0, 32, h->stream) != cudaSuccess)
    return set_error(CDAE_E_CUDA, "radix sort (size query) failed");
  TRY(ensure(h, h->tc_probe_tmp, std::max<size_t>(tmp_bytes, 16)));
  {
    ProfScope ps(h, CDAE_K_TOPN_PACK);
    CU(cudaMemsetAsync(h->tc_probe_zsum.p, 0, sizeof(float) * h->ld, h->stream));
    const int upb = 64;
    tc::probe_colsum_kernel<<<cdiv(n_users, upb), 256, 0, h->stream>>>(h->topn_z.p, users, (int)n_users, h->ld, upb, h->tc_probe_zsum.p);
    KERNEL_OK(h);
    tc::probe_key_kernel<<<cdiv(h->I * 32, 256), 256, 0, h->stream>>>(Wd, h->m.bp, h->I, K, h->ld, h->tc_probe_zsum.p,
                                                                      1.f / (float)n_users, keys_in, ids_in);
    KERNEL_OK(h);
    tmp_bytes = h->tc_probe_tmp.cap;
    if (cub::DeviceRadixSort::SortPairsDescending(h->tc_probe_tmp.p, tmp_bytes, keys_in, keys_out, ids_in, ids_out, I, 0, 32, h->stream) != cudaSuccess)
      return set_error(CDAE_E_CUDA, "radix sort failed");
    CU(cudaMemsetAsync(h->tc_probe_pos.p, 0xff, sizeof(int32_t) * h->I, h->stream));
    tc::probe_pos_kernel<<<cdiv(M, 256), 256, 0, h->stream>>>(ids_out, M, h->tc_probe_pos.p);
    KERNEL_OK(h);
    tc::probe_gather_w_kernel<<<cdiv((int64_t)M * (Kp / 8), 256), 256, 0, h->stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(h->tc_wb.p), ids_out, M, Kp, reinterpret_cast<__nv_bfloat16*>(h->tc_probe_wb.p));
    KERNEL_OK(h);
    CU(cudaMemsetAsync(h->tc_probe_bits.p, 0, sizeof(uint32_t) * (size_t)(n_pad * words), h->stream));
    tc::probe_bitmap_kernel<<<cdiv(n_users * 32, 256), 256, 0, h->stream>>>(users, (int)n_users, h->row_ptr_d.p, h->col_d.p,
                                                                           h->tc_probe_pos.p, words, h->tc_probe_bits.p);
    KERNEL_OK(h);
    tc::pack_z_bf16_kernel<<<cdiv(n_pad * 32, 256), 256, 0, h->stream>>>(
        h->topn_z.p, users, (int)n_users, n_pad, K, h->ld, Kp, h->tc_wmax.p,
        reinterpret_cast<__nv_bfloat16*>(h->tc_zb.p), h->tc_eps.p);
    KERNEL_OK(h);
  }
  alignas(64) CUtensorMap ma, mb;
  TRY(tc_make_map(&ma, h->tc_zb.p, (uint64_t)n_pad, (uint64_t)Kp, tc::TILE_U));
  TRY(tc_make_map(&mb, h->tc_probe_wb.p, (uint64_t)M, (uint64_t)Kp, tc::TILE_I));
  tc::TcArgs a;
  a.n_users = (int)n_users; a.I = M; a.n_tiles = M / tc::TILE_I;
  a.users = users; a.row_ptr = h->row_ptr_d.p; a.col = h->col_d.p;
  a.cand_id = h->cand_id.p; a.cand_s = h->cand_s.p; a.cand_cnt = h->cand_cnt.p; a.cand_thr = h->tc_thr.p;
  a.init_thr = nullptr;
  a.bits = h->tc_probe_bits.p;
  a.words = words;
  a.n_splits = 1;
  a.tiles_per_split = a.n_tiles;
  a.seg = tc::CAND_MAX;
  const int grid = (int)(n_pad / tc::TILE_U);
  {
    ProfScope ps(h, CDAE_K_TOPN);
    switch (KB) {
      case 1: TRY(tc_launch<1>(h, ma, mb, a, grid)); break;
      case 2: TRY(tc_launch<2>(h, ma, mb, a, grid)); break;
      case 3: TRY(tc_launch<3>(h, ma, mb, a, grid)); break;
      case 4: TRY(tc_launch<4>(h, ma, mb, a, grid)); break;
      default: TRY(tc_launch<5>(h, ma, mb, a, grid)); break;
    }
    KERNEL_OK(h);
  }
  {
    ProfScope ps(h, CDAE_K_TOPN_PACK);
    tc::probe_thr_kernel<<<cdiv(n_users, 256), 256, 0, h->stream>>>(h->cand_s.p, h->cand_cnt.p, h->tc_eps.p, (int)n_users,
                                                                    a.seg, topk, h->tc_probe_thr.p);
    KERNEL_OK(h);
  }
  *m_out = M;
  return 0;
}

// Tensor-core candidate phase: sweep 1 over all users from thr = -inf; sweep 2 over the (few)
// users it could not prove, starting at the threshold sweep 1 derived for each of them, which
// normally leaves only a handful of items above it; what is still unproven goes to the exact
// fp32 kernel (*n_redo users, ids in h->tc_redo).
static int topn_candidates_tc(cdae_handle* h, const float* Wd, const int32_t* users, int64_t n_users,
                              int topk, int* n_redo) {
  const int K = h->K, Kp = (int)round_up(K + 2, tc::KBLK);
  const int64_t I_pad = round_up(h->I, tc::TILE_I);
  TRY(ensure(h, h->tc_wb, (size_t)(I_pad * Kp)));
  TRY(ensure(h, h->tc_wmax, (size_t)round_up(K + 2, 4)));
  TRY(ensure(h, h->tc_eps, (size_t)n_users));
  TRY(ensure(h, h->tc_thr, (size_t)n_users));
  // [ids sweep 1 | ids sweep 2 | counter] and the matching start thresholds
  TRY(ensure(h, h->tc_redo, (size_t)(2 * n_users + 4)));
  TRY(ensure(h, h->tc_redo_thr, (size_t)n_users));
  int32_t* redo1 = h->tc_redo.p;
  int32_t* redo2 = h->tc_redo.p + n_users;
  int* redo_cnt = h->tc_redo.p + 2 * n_users;
  CU(cudaMemsetAsync(h->tc_wmax.p, 0, sizeof(float) * (K + 2), h->stream));
  {
    ProfScope ps(h, CDAE_K_TOPN_PACK);
    tc::absmax_cols_kernel<<<std::min<int>(cdiv(h->I, 128), h->sm_count * 4), 256, 0, h->stream>>>(
        Wd, h->m.bp, h->I, K, h->ld, h->tc_wmax.p);
    KERNEL_OK(h);
    tc::rownorm_max_kernel<<<cdiv(h->I * 32, 256), 256, 0, h->stream>>>(Wd, h->I, K, h->ld, h->tc_wmax.p);
    KERNEL_OK(h);
    tc::pack_w_bf16_kernel<<<cdiv(I_pad * (Kp / 8), 256), 256, 0, h->stream>>>(
        Wd, h->m.bp, h->I, I_pad, K, h->ld, Kp, reinterpret_cast<__nv_bfloat16*>(h->tc_wb.p));
    KERNEL_OK(h);
  }
  int n1 = 0, probe_m = 0;
  TRY(tc_probe_thresholds(h, Wd, users, n_users, topk, &probe_m));
  h->topn_probe_items = probe_m;
  TRY(tc_pass(h, Wd, users, n_users, topk, probe_m > 0 ? h->tc_probe_thr.p : nullptr, redo1, h->tc_redo_thr.p, redo_cnt, &n1,
              /*allow_split=*/false));
  h->topn_pass2_users = n1;
  *n_redo = n1;
  h->tc_exact_list = redo1;
  if (n1 > 0) {
    int n2 = 0;
    TRY(tc_pass(h, Wd, redo1, n1, topk, h->tc_redo_thr.p, redo2, nullptr, redo_cnt, &n2));
    *n_redo = n2;
    h->tc_exact_list = redo2;
  }
  return 0;
}

extern "C" {

int cdae_topn_build(cdae_handle* h, int32_t topk) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  if (topk < 1 || topk > TOPN_MAX_K) return set_error(CDAE_E_INVALID, "topk must be in [1,%d]", TOPN_MAX_K);
  TRY(begin_call(h));
  TRY(encode_all_users(h));
  const int32_t* users = h->world > 1 ? h->plan_uids.p : nullptr;
  int64_t n_users = h->U;
  if (h->world > 1) {
    n_users = 0;
    for (const MiniBatch& p : h->plan) n_users += p.n_users;
  }
  const int64_t slots = std::max<int>(TOPN_M, tc::CAND_MAX);
  TRY(ensure(h, h->cand_id, (size_t)(n_users * slots)));
  TRY(ensure(h, h->cand_s, (size_t)(n_users * slots)));
  TRY(ensure(h, h->cand_cnt, (size_t)n_users));
  TRY(ensure(h, h->flag_d, 4));
  TRY(ensure(h, h->topn_ids, (size_t)(h->U * topk)));
  TRY(ensure(h, h->topn_scores, (size_t)(h->U * topk)));
  CU(cudaMemsetAsync(h->topn_ids.p, 0xff, sizeof(int32_t) * h->U * topk, h->stream));
  CU(cudaMemsetAsync(h->topn_scores.p, 0, sizeof(float) * h->U * topk, h->stream));
  CU(cudaMemsetAsync(h->flag_d.p, 0, sizeof(int) * 4, h->stream));
  const float* Wd = h->m.asym ? h->m.V : h->m.W;
  h->topn_tc_users = h->topn_redo_users = 0;
  h->topn_path = 0;
  h->topn_probe_items = 0;
  const int32_t* exact_users = users;
  int64_t n_exact = n_users;
  if (n_users > 0 && tc_path_wanted(h, topk)) {
    int n_redo = 0;
    TRY(topn_candidates_tc(h, Wd, users, n_users, topk, &n_redo));
    h->topn_path = 1;
    h->topn_tc_users = n_users - n_redo;
    h->topn_redo_users = n_redo;
    exact_users = h->tc_exact_list;
    n_exact = n_redo;
  }
  if (n_exact > 0) {
    TRY(topn_candidates(h, Wd, exact_users, n_exact));
    ProfScope ps(h, CDAE_K_TOPN_RERANK);
    topn_rerank_kernel<<<cdiv(n_exact, 8), 256, 0, h->stream>>>(
        h->topn_z.p, Wd, h->m.bp, h->K, h->ld, exact_users, (int)n_exact, h->cand_id.p, h->cand_cnt.p, TOPN_M, 1,
        topk, h->topn_ids.p, h->topn_scores.p, h->flag_d.p, nullptr, nullptr, nullptr, nullptr, nullptr);
    KERNEL_OK(h);
  }
  h->topn_ids_h.resize((size_t)(h->U * topk));
  h->topn_scores_h.resize((size_t)(h->U * topk));
  int flag[4] = {0, 0, 0, 0};
  CU(cudaMemcpyAsync(h->topn_ids_h.data(), h->topn_ids.p, sizeof(int32_t) * h->U * topk, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(h->topn_scores_h.data(), h->topn_scores.p, sizeof(float) * h->U * topk, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(flag, h->flag_d.p, sizeof(flag), cudaMemcpyDeviceToHost, h->stream));
  h->d2h += (sizeof(int32_t) + sizeof(float)) * h->U * topk;
  TRY(end_call(h, nullptr));
  h->topn_k = topk;
  if (flag[0])
    return set_error(CDAE_E_INVALID, "a user has fewer than topk unrated items "
                     "(the reference CHECK-aborts here, cdae.hpp:187)");
  return 0;
}

/* Which candidate path the last cdae_topn_build used (1 = tcgen05, 0 = fp32 CUDA cores) and how
 * many users its error bound verified / how many were redone by the exact kernel. */
int cdae_topn_stats(cdae_handle* h, int32_t* path, int64_t* verified_users, int64_t* redone_users) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  if (path) *path = h->topn_path;
  if (verified_users) *verified_users = h->topn_tc_users;
  if (redone_users) *redone_users = h->topn_redo_users;
  return 0;
}

int cdae_topn_probe_items(cdae_handle* h, int32_t* items_out) {
  if (!h || !items_out) return set_error(CDAE_E_INVALID, "NULL argument");
  *items_out = h->topn_probe_items;
  return 0;
}

int cdae_topn_lookup(cdae_handle* h, int64_t uid, int64_t* ids_out, float* scores_out) {
  if (!h || !ids_out) return set_error(CDAE_E_INVALID, "NULL argument");
  if (h->topn_k == 0) return set_error(CDAE_E_STATE, "cdae_topn_build has not run");
  if (uid < 0 || uid >= h->U) return set_error(CDAE_E_INVALID, "uid out of range");
  const int k = h->topn_k;
  for (int t = 0; t < k; ++t) {
    ids_out[t] = h->topn_ids_h[(size_t)(uid * k + t)];
    if (scores_out) scores_out[t] = h->topn_scores_h[(size_t)(uid * k + t)];
  }
  return 0;
}

int cdae_topn_fetch(cdae_handle* h, int64_t* ids_out, float* scores_out) {
  if (!h || !ids_out) return set_error(CDAE_E_INVALID, "NULL argument");
  if (h->topn_k == 0) return set_error(CDAE_E_STATE, "cdae_topn_build has not run");
  const size_t n = (size_t)(h->U * h->topn_k);
  for (size_t i = 0; i < n; ++i) ids_out[i] = h->topn_ids_h[i];
  if (scores_out) memcpy(scores_out, h->topn_scores_h.data(), sizeof(float) * n);
  return 0;
}

int cdae_topn_evaluate(cdae_handle* h, const int64_t* trp, const int32_t* tcol, double* out8,
                       int64_t* users_evaluated) {
  if (!h || !trp || !out8) return set_error(CDAE_E_INVALID, "NULL argument");
  if (h->topn_k == 0) return set_error(CDAE_E_STATE, "cdae_topn_build has not run");
  if (h->world > 1) return set_error(CDAE_E_STATE, "evaluate per rank with cdae_topn_fetch in a process group");
  CU(cudaSetDevice(h->cfg.device));
  const int64_t nnz = trp[h->U];
  if (nnz > 0 && !tcol) return set_error(CDAE_E_INVALID, "test_col is NULL");
  TRY(validate_csr(h->U, h->I, trp, tcol));  // topn_metrics_kernel binary-searches the test rows
  TRY(ensure(h, h->test_rp_d, (size_t)h->U + 1));
  TRY(ensure(h, h->test_col_d, (size_t)std::max<int64_t>(nnz, 1)));
  TRY(ensure(h, h->stage_d, 9));
  CU(cudaMemcpyAsync(h->test_rp_d.p, trp, sizeof(int64_t) * (h->U + 1), cudaMemcpyHostToDevice, h->stream));
  if (nnz) CU(cudaMemcpyAsync(h->test_col_d.p, tcol, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(h->stage_d.p, 0, sizeof(double) * 9, h->stream));
  topn_metrics_kernel<<<cdiv(h->U, 256), 256, 0, h->stream>>>(h->topn_ids.p, h->topn_k, h->U,
                                                             h->test_rp_d.p, h->test_col_d.p, h->stage_d.p);
  KERNEL_OK(h);
  double r[9];
  CU(cudaMemcpyAsync(r, h->stage_d.p, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  // mean over users that have test items (evaluation.hpp:160-166)
  for (int i = 0; i < 8; ++i) out8[i] = r[8] > 0 ? r[i] / r[8] : 0.;
  if (users_evaluated) *users_evaluated = (int64_t)r[8];
  return 0;
}

}  // extern "C"
