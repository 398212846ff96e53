// cdae_b200/csrc/fulldec_api.inl — host side of full-item-decode training (fulldec_tc.cuh),
// included at the end of api.cu.  One frozen minibatch:
//   pack (W' + b' -> Wb, Z -> Zb, targets -> bitmap)  ->  fd_score (G)  ->  fd_gemm<hidden> (HG)  ->  fd_gemm<itemgrad> (gW', gb')
//   [CDAE_B200_FD=fused: fd_fused (G, HG) -> fd_gemm<itemgrad>]
// in place of decode_kernel; everything before (gather, activate) and after (hidden_backward,
// scatter, all-reduce, apply) is the sampled path's.

// How many CTAs share the `units` pieces of work of each of `tiles` output tiles: wave efficiency of a
// split S is tiles*S / (ceil(tiles*S / SMs) * SMs); the smallest S within 8 % of the best one wins.
static int fd_pick_split(int tiles, int units, int min_units, int sms, int max_s) {
  double eff[64];
  double best_eff = 0.;
  int n = 0;
  for (int s = 1; s <= max_s && s < 64; ++s) {
    if (s > 1 && units / s < min_units) break;
    const int ctas = tiles * s;
    eff[s] = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
    best_eff = std::max(best_eff, eff[s]);
    n = s;
  }
  // every extra split re-reads one operand and adds a pass of output reductions: take the
  // smallest split that is within 8% of the best wave efficiency
  for (int s = 1; s <= n; ++s)
    if (eff[s] >= 0.92 * best_eff) return s;
  return 1;
}

template <int KB, int LT, bool BOUT>
static int fd_launch_score(cdae_handle* h, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb_half,
                           const CUtensorMap& mg, const fd::ScoreArgs& a, dim3 grid, bool cluster2) {
  const size_t dyn = fd::score_smem(KB);
  {  // per launch: the attribute belongs to the CURRENT device (several handles / devices per process)
    CU(cudaFuncSetAttribute(fd::fd_score_kernel<KB, LT, BOUT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    CU(cudaFuncSetAttribute(fd::fd_score_kernel<KB, LT, BOUT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  }
  if (!cluster2) {
    fd::fd_score_kernel<KB, LT, BOUT, false><<<grid, 640, dyn, h->stream>>>(ma, mb, mg, a);
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = dyn; cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CU(cudaLaunchKernelEx(&cfg, fd::fd_score_kernel<KB, LT, BOUT, true>, ma, mb_half, mg, a));
  return 0;
}
template <int KB, bool ITEMGRAD, bool BOUT>
static int fd_launch_gemm(cdae_handle* h, const CUtensorMap& ma, const CUtensorMap& mb, const fd::GemmArgs& a, dim3 grid) {
  const size_t dyn = fd::gemm_smem(KB);
  {  // per launch: the attribute belongs to the CURRENT device (several handles / devices per process)
    CU(cudaFuncSetAttribute(fd::fd_gemm_kernel<KB, ITEMGRAD, BOUT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    CU(cudaFuncSetAttribute(fd::fd_gemm_kernel<KB, ITEMGRAD, BOUT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  }
  // CDAE_B200_FD_GEMM_CLUSTER=1: pairs of output tiles share the B stream (Wb / Zb slabs of each
  // contraction step) through 2-CTA clusters with multicast TMA.  Parity-identical, measured neutral
  // (hidden 0.412 vs 0.404 ms, item gradient 0.484 vs 0.482 ms per two launches at config C): both
  // kernels are bound by reading G from HBM (62-68 % of the nominal peak = ~80 % of the measured copy
  // bandwidth, profiles/r01_q_*), not by the B stream out of L2.
  static const bool cl_on = getenv("CDAE_B200_FD_GEMM_CLUSTER") && atoi(getenv("CDAE_B200_FD_GEMM_CLUSTER")) != 0;
  if (!cl_on || grid.x % 2 != 0) {
    fd::fd_gemm_kernel<KB, ITEMGRAD, BOUT, false><<<grid, 256, dyn, h->stream>>>(ma, mb, a);
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = dyn; cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CU(cudaLaunchKernelEx(&cfg, fd::fd_gemm_kernel<KB, ITEMGRAD, BOUT, true>, ma, mb, a));
  return 0;
}
template <int KB, int LT>
static int fd_launch_fused(cdae_handle* h, const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mw_half,
                           const CUtensorMap& mg, const fd::FusedArgs& a, dim3 grid, bool cluster2) {
  const size_t dyn = fd::fused_smem(KB);
  {  // per launch: the attribute belongs to the CURRENT device (several handles / devices per process)
    CU(cudaFuncSetAttribute(fd::fd_fused_kernel<KB, LT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    CU(cudaFuncSetAttribute(fd::fd_fused_kernel<KB, LT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  }
  if (!cluster2) {
    fd::fd_fused_kernel<KB, LT, false><<<grid, 640, dyn, h->stream>>>(ma, mw, mg, a);
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = dyn; cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CU(cudaLaunchKernelEx(&cfg, fd::fd_fused_kernel<KB, LT, true>, ma, mw_half, mg, a));
  return 0;
}
#define FD_DISPATCH_KB(KBV, CALL)      \
  switch (KBV) {                       \
    case 1: TRY(CALL(1)); break;       \
    case 2: TRY(CALL(2)); break;       \
    case 3: TRY(CALL(3)); break;       \
    default: TRY(CALL(4)); break;      \
  }

static bool K_has_bias_room(int K) { return K + 2 <= (int)round_up(K, tc::KBLK); }

// H5-H7 with the output set = all items: fills bt.HG and adds the decoder-side
// gradients of this rank's users to gW' / gb'.  bt.Z must hold the hidden activations.
static int run_fulldec(cdae_handle* h, const BatchDev& bt) {
  if (bt.n_users == 0) return 0;
  // the output bias rides in the contraction (two spare columns) unless K leaves no room for them
  const bool bias_in = K_has_bias_room(h->K);
  const int K = h->K, Kp = (int)round_up(bias_in ? K + 2 : K, tc::KBLK), KB = Kp / tc::KBLK;
  const int64_t B_pad = round_up(bt.n_users, 128), I_pad = round_up(h->I, tc::TILE_I);
  if ((double)B_pad * (double)I_pad * 2.0 > 64e9)
    return set_error(CDAE_E_INVALID, "full_decode: the loss-gradient matrix of one minibatch (%lld users x %lld items, bf16) "
                     "would take more than 64 GB; lower batch_users", (long long)B_pad, (long long)I_pad);
  TRY(ensure(h, h->fd_zb, (size_t)(B_pad * Kp)));
  TRY(ensure(h, h->fd_wb, (size_t)(I_pad * Kp)));
  TRY(ensure(h, h->fd_g, (size_t)(B_pad * I_pad)));
  const int64_t words = I_pad / 32;
  TRY(ensure(h, h->fd_bits, (size_t)(B_pad * words)));
  const float* Wd = h->m.asym ? h->m.V : h->m.W;
  float* gWd = h->m.asym ? h->m.gV : h->m.gW;
  __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(h->fd_zb.p);
  __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(h->fd_wb.p);
  __nv_bfloat16* G = reinterpret_cast<__nv_bfloat16*>(h->fd_g.p);
  {
    ProfScope ps(h, CDAE_K_FD_PACK);
    if (bias_in) {
      tc::pack_w_bf16_kernel<<<cdiv(I_pad * (Kp / 8), 256), 256, 0, h->stream>>>(Wd, h->m.bp, h->I, I_pad, K, h->ld, Kp, wb);
    } else {
      TRY(ensure(h, h->fd_bias, (size_t)I_pad));
      fd::pack_w_plain_kernel<<<cdiv(I_pad * (Kp / 8), 256), 256, 0, h->stream>>>(Wd, h->m.bp, h->I, I_pad, K, h->ld, Kp, wb, h->fd_bias.p);
    }
    KERNEL_OK(h);
    CU(cudaMemsetAsync(h->fd_bits.p, 0, sizeof(uint32_t) * (size_t)(B_pad * words), h->stream));
    fd::fd_bitmap_kernel<<<cdiv((int64_t)bt.n_users * 32, 256), 256, 0, h->stream>>>(bt.uids, bt.n_users, bt.row_ptr, bt.col, words, h->fd_bits.p);
    KERNEL_OK(h);
    fd::pack_z_train_kernel<<<cdiv(B_pad * (Kp / 8), 256), 256, 0, h->stream>>>(bt.Z, bt.n_users, B_pad, K, h->ld, Kp, bias_in ? 1 : 0, zb);
    KERNEL_OK(h);
  }
  alignas(64) CUtensorMap m_zb_a, m_wb_b, m_g_st, m_g_rows, m_g_cols, m_wb_mn, m_zb_mn;
  TRY(tc_make_map(&m_g_st, G, (uint64_t)B_pad, (uint64_t)I_pad, 32));          // score: G store, {64 items, 32 users}
  TRY(tc_make_map(&m_zb_a, zb, (uint64_t)B_pad, (uint64_t)Kp, tc::TILE_U));    // score: A = Zb, 128-user boxes
  TRY(tc_make_map(&m_wb_b, wb, (uint64_t)I_pad, (uint64_t)Kp, tc::TILE_I));    // score: B = Wb, 256-item boxes
  TRY(tc_make_map(&m_g_rows, G, (uint64_t)B_pad, (uint64_t)I_pad, 128));       // hidden: A = G, {64 items, 128 users}
  TRY(tc_make_map(&m_g_cols, G, (uint64_t)B_pad, (uint64_t)I_pad, 64));        // itemgrad: A = G^T, {64 items, 64 users}
  TRY(tc_make_map(&m_wb_mn, wb, (uint64_t)I_pad, (uint64_t)Kp, 64));           // hidden: B = Wb, {64 cols, 64 items}
  TRY(tc_make_map(&m_zb_mn, zb, (uint64_t)B_pad, (uint64_t)Kp, 64));           // itemgrad: B = Zb, {64 cols, 64 users}
  const int u_tiles = (int)(B_pad / 128);
  // CDAE_B200_FD=fused: score + hidden-gradient contraction in ONE kernel (fd_fused_kernel).  Parity-
  // identical; measured 9 % slower than the two launches at config C (profiles/r01_m_*).  What was
  // tried: 2-CTA multicast of the W' tiles (halves the L2 stream: no change), a second W' ring so
  // that no slot is held across the epilogue (+32 KB of shared-memory writes per tile: 12 % SLOWER).
  // Both say the kernel is bound by shared-memory bandwidth: with 64-item tiles every N = 64 MMA
  // re-reads its 4 KB A slice for 2 KB of B, and the second contraction adds 12 KB per k-step.
  // The split path is the default.
  static const bool fused_path = getenv("CDAE_B200_FD") && strcmp(getenv("CDAE_B200_FD"), "fused") == 0;
  if (fused_path && bias_in) {
    fd::FusedArgs a;
    a.n_users = bt.n_users; a.I = h->I; a.I_pad = I_pad; a.n_tiles = (int)(I_pad / fd::FU_TILE_I);
    a.ksteps = (K + 2 + 15) / 16; a.K = K; a.ld = h->ld;
    a.bits = h->fd_bits.p; a.HG = bt.HG;
    a.outputs = &h->stats_d->outputs[0];
    const int S = fd_pick_split(u_tiles, a.n_tiles, 16, h->sm_count, 32);
    a.tiles_per_split = (a.n_tiles + S - 1) / S;
    const dim3 grid(u_tiles, (a.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split);
    static const bool cl_off = getenv("CDAE_B200_FD_CLUSTER") && atoi(getenv("CDAE_B200_FD_CLUSTER")) == 0;
    const bool cluster2 = !cl_off && (u_tiles % 2 == 0);
    alignas(64) CUtensorMap m_wb_q;
    TRY(tc_make_map(&m_wb_q, wb, (uint64_t)I_pad, (uint64_t)Kp, fd::FU_TILE_I / 2));
    ProfScope ps(h, CDAE_K_FD_SCORE);
    if (h->m.loss == LOSS_CE) {
#define CALL(KBV) fd_launch_fused<KBV, LOSS_CE>(h, m_zb_a, m_wb_mn, m_wb_q, m_g_rows, a, grid, cluster2)
      FD_DISPATCH_KB(KB, CALL)
#undef CALL
    } else {
#define CALL(KBV) fd_launch_fused<KBV, LOSS_SQUARE>(h, m_zb_a, m_wb_mn, m_wb_q, m_g_rows, a, grid, cluster2)
      FD_DISPATCH_KB(KB, CALL)
#undef CALL
    }
    KERNEL_OK(h);
  } else {
  {
    fd::ScoreArgs a;
    a.n_users = bt.n_users; a.I = h->I; a.I_pad = I_pad; a.n_tiles = (int)(I_pad / tc::TILE_I);
    a.ksteps = ((bias_in ? K + 2 : K) + 15) / 16;
    a.bits = h->fd_bits.p; a.G = G;
    a.bias = bias_in ? nullptr : h->fd_bias.p;
    a.outputs = &h->stats_d->outputs[0];
    const int S = fd_pick_split(u_tiles, a.n_tiles, 4, h->sm_count, 32);
    a.tiles_per_split = (a.n_tiles + S - 1) / S;
    const dim3 grid(u_tiles, (a.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split);
    // CDAE_B200_FD_CLUSTER=1: pairs of user tiles share their W' stream through 2-CTA clusters
    // (parity-identical, measured neutral: see fd_score_kernel)
    static const bool cl_on = getenv("CDAE_B200_FD_CLUSTER") && atoi(getenv("CDAE_B200_FD_CLUSTER")) != 0;
    const bool cluster2 = cl_on && (u_tiles % 2 == 0);
    alignas(64) CUtensorMap m_wb_half;
    TRY(tc_make_map(&m_wb_half, wb, (uint64_t)I_pad, (uint64_t)Kp, tc::TILE_I / 2));
    ProfScope ps(h, CDAE_K_FD_SCORE);
    if (h->m.loss == LOSS_CE && bias_in) {
#define CALL(KBV) fd_launch_score<KBV, LOSS_CE, false>(h, m_zb_a, m_wb_b, m_wb_half, m_g_st, a, grid, cluster2)
      FD_DISPATCH_KB(KB, CALL)
#undef CALL
    } else if (h->m.loss == LOSS_CE) {
#define CALL(KBV) fd_launch_score<KBV, LOSS_CE, true>(h, m_zb_a, m_wb_b, m_wb_half, m_g_st, a, grid, cluster2)
      FD_DISPATCH_KB(KB, CALL)
#undef CALL
    } else if (bias_in) {
#define CALL(KBV) fd_launch_score<KBV, LOSS_SQUARE, false>(h, m_zb_a, m_wb_b, m_wb_half, m_g_st, a, grid, cluster2)
      FD_DISPATCH_KB(KB, CALL)
#undef CALL
    } else {
#define CALL(KBV) fd_launch_score<KBV, LOSS_SQUARE, true>(h, m_zb_a, m_wb_b, m_wb_half, m_g_st, a, grid, cluster2)
      FD_DISPATCH_KB(KB, CALL)
#undef CALL
    }
    KERNEL_OK(h);
  }
  {
    fd::GemmArgs a{};
    a.n_steps = (int)(I_pad / 64);
    const int S = fd_pick_split(u_tiles, a.n_steps, 8, h->sm_count, 16);
    a.steps_per_split = (a.n_steps + S - 1) / S;
    a.n_rows = bt.n_users; a.K = K; a.ld = h->ld;
    a.out = bt.HG; a.out_bias = nullptr; a.W = nullptr; a.bp = nullptr; a.nlambda = 0.f;
    const dim3 grid(u_tiles, (a.n_steps + a.steps_per_split - 1) / a.steps_per_split);
    ProfScope ps(h, CDAE_K_FD_HIDDEN);
#define CALL(KBV) fd_launch_gemm<KBV, false, false>(h, m_g_rows, m_wb_mn, a, grid)
    FD_DISPATCH_KB(KB, CALL)
#undef CALL
    KERNEL_OK(h);
  }
  }
  {
    fd::GemmArgs a{};
    a.n_steps = (int)(B_pad / 64);
    const int i_tiles = (int)(I_pad / 128);
    const int S = fd_pick_split(i_tiles, a.n_steps, 8, h->sm_count, 16);
    a.steps_per_split = (a.n_steps + S - 1) / S;
    a.n_rows = (int)h->I; a.K = K; a.ld = h->ld;
    a.out = gWd; a.out_bias = h->m.gbp; a.W = Wd; a.bp = h->m.bp;
    a.nlambda = (float)bt.n_users * h->m.lambda;
    const dim3 grid(i_tiles, (a.n_steps + a.steps_per_split - 1) / a.steps_per_split);
    ProfScope ps(h, CDAE_K_FD_ITEMGRAD);
    if (bias_in) {
#define CALL(KBV) fd_launch_gemm<KBV, true, false>(h, m_g_cols, m_zb_mn, a, grid)
      FD_DISPATCH_KB(KB, CALL)
#undef CALL
    } else {
#define CALL(KBV) fd_launch_gemm<KBV, true, true>(h, m_g_cols, m_zb_mn, a, grid)
      FD_DISPATCH_KB(KB, CALL)
#undef CALL
    }
    KERNEL_OK(h);
  }
  return 0;
}
