// cdae_b200/csrc/topn_api.inl — cdae_topn_* entry points (included at the end of api.cu).

// Uncorrupted hidden vectors of this rank's users into topn_z (cdae.hpp:167-172: scale = 1;
// with corruption_ratio == 1 the reference feeds an EMPTY input set).
static int encode_all_users(cdae_handle* h) {
  if (!h->plan_valid) TRY(build_plan(h));
  TRY(ensure_scratch(h, h->plan_max_users, h->plan_max_slots));
  TRY(ensure(h, h->topn_z, (size_t)(h->U * h->ld)));
  const size_t per = (size_t)std::max<int64_t>(h->scratch_users, 1) * h->ld;
  const bool empty_input = h->cfg.corruption_ratio == 1.;
  for (const MiniBatch& p : h->plan) {
    if (p.n_users == 0) continue;
    BatchDev bt = make_batch(h, h->plan_in.p + p.in0, p.n_in, h->plan_out.p + p.out0, 0,
                             h->plan_uids.p + p.user0, p.n_users);
    bt.Z = h->topn_z.p + p.uid0 * h->ld;  // a slice's users are consecutive ids
    CU(cudaMemsetAsync(h->keep.p, empty_input ? 0 : 1, (size_t)std::max<int64_t>(p.slots, 1), h->stream));
    CU(cudaMemsetAsync(h->acc3.p, 0, sizeof(float) * per, h->stream));
    TRY(launch_gather(h, bt, nullptr, false));
    TRY(launch_activate(h, bt, 1.f));
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
  TRY(ensure(h, h->cand_id, (size_t)(n_users * TOPN_M)));
  TRY(ensure(h, h->cand_s, (size_t)(n_users * TOPN_M)));
  TRY(ensure(h, h->cand_cnt, (size_t)n_users));
  TRY(ensure(h, h->flag_d, 4));
  TRY(ensure(h, h->topn_ids, (size_t)(h->U * topk)));
  TRY(ensure(h, h->topn_scores, (size_t)(h->U * topk)));
  CU(cudaMemsetAsync(h->topn_ids.p, 0xff, sizeof(int32_t) * h->U * topk, h->stream));
  CU(cudaMemsetAsync(h->topn_scores.p, 0, sizeof(float) * h->U * topk, h->stream));
  CU(cudaMemsetAsync(h->flag_d.p, 0, sizeof(int) * 4, h->stream));
  const float* Wd = h->m.asym ? h->m.V : h->m.W;
  TRY(topn_candidates(h, Wd, users, n_users));
  topn_rerank_kernel<<<cdiv(n_users, 8), 256, 0, h->stream>>>(
      h->topn_z.p, Wd, h->m.bp, h->K, h->ld, users, (int)n_users, h->cand_id.p, h->cand_cnt.p, topk,
      h->topn_ids.p, h->topn_scores.p, h->flag_d.p);
  KERNEL_OK(h);
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
