// cdae_b200/csrc/handle.cuh — the state behind an opaque cdae_handle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/cdae_b200.h"
#include "train_kernels.cuh"

namespace cdae {

int set_error(int code, const char* fmt, ...);

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// This rank's slice of one global minibatch, as ranges into the plan_* lists.
struct MiniBatch {
  int64_t user0, n_users;  // range in plan_uids
  int64_t uid0;            // first global uid of the slice (slices are contiguous)
  int64_t in0, n_in;       // range in plan_in
  int64_t out0, n_out;     // range in plan_out
  int64_t slots;           // train items of the slice (size of keep[], negs[] / num_neg)
};

}  // namespace cdae

struct cdae_handle {
  cdae_config_t cfg;
  int64_t U = 0, I = 0, I4 = 0, nnz = 0;
  int K = 0, ld = 0;
  int64_t batch_users = 0;
  int ch_in = 64, ch_out = 16;
  int rank = 0, world = 1;
  void* comm = nullptr;  // ncclComm_t
  // NVLink peer-memory all-reduce (p2p_allreduce.cuh), opt-in through cdae_dist_p2p_open
  bool p2p_on = false;
  bool p2p_fused = false;          // reduce-scatter + sliced optimiser step + all-gather in one kernel (default)
  bool p2p_ipc = false;            // peer pointers came from cudaIpcOpenMemHandle (other processes)
  float* p2p_bufs[8] = {nullptr};  // every rank's gradient buffer(s): base of [2][grad_floats] when fused
  float* p2p_params[8] = {nullptr};  // every rank's item-side parameter buffer
  uint32_t* p2p_flags[8] = {nullptr};
  uint32_t* p2p_my_flags = nullptr;
  unsigned int* p2p_done = nullptr;
  unsigned long long* p2p_ts = nullptr;   // phase timestamps of the fused combine kernel (cdae_debug_combine only)
  int p2p_parity = 0;              // which gradient buffer the current minibatch accumulates into
  // NVLS mode (mc_nvls.inl): the item side lives in a VMM block bound to a multicast object
  bool mc_creator = false, mc_attached = false, mc_active = false;
  bool mc_shared = false;          // the multicast object handle belongs to another handle of this process (cdae_group)
  unsigned long long mc_handle = 0, mc_phys = 0;   // CUmemGenericAllocationHandle
  size_t mc_size = 0;
  char* mc_uc = nullptr;           // this rank's block (unicast mapping): [flags 4 KB | parameters | gradients x2]
  char* mc_mc = nullptr;           // the same offsets through the multicast mapping
  uint32_t p2p_epoch = 0;
  int sm_count = 148;
  size_t dev_bytes = 0;

  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;      // cdae_train_epoch_csr: host -> device pieces, one event per minibatch
  std::vector<cudaEvent_t> copy_ev;
  cudaStream_t side_stream = nullptr;      // hidden_backward_kernel runs here, beside scatter_kernel on `stream`
  cudaEvent_t side_fork = nullptr, side_join = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cdae::ModelDev m;
  // item side: parameters, AdaGrad state and minibatch gradients as flat buffers of ONE layout
  // (point_item_side in api.cu); h->m.W / V / bp / b and their *_ag / g* pointers point into them
  cdae::DevBuf<float> item_params, item_acc;
  cdae::DevBuf<float> grad;      // [grad_floats], or 2 x grad_floats (ping-pong) in peer-memory mode
  size_t grad_floats = 0;

  std::vector<int64_t> row_ptr_h;  // host copy (work lists depend on it)
  cdae::DevBuf<int64_t> row_ptr_d;
  cdae::DevBuf<int32_t> col_d;

  // epoch plan
  std::vector<cdae::MiniBatch> plan;
  cdae::DevBuf<cdae::WorkItem> plan_in, plan_out;
  cdae::DevBuf<int32_t> plan_uids;
  int64_t plan_max_slots = 0, plan_max_users = 0;
  size_t plan_n_uids = 0;        // users this rank trains per epoch (length of plan_uids)
  bool plan_valid = false;
  bool csr_bad = false;          // ... and failed: training calls are refused until a valid CSR arrives
  bool csr_unchecked = false;    // the device CSR came from cdae_train_epoch_csr and has not been validated yet
  // explicit-user calls
  cdae::DevBuf<cdae::WorkItem> tmp_in, tmp_out;
  cdae::DevBuf<int32_t> tmp_uids;

  // per-minibatch scratch
  cdae::DevBuf<uint8_t> keep;
  cdae::DevBuf<int32_t> negs;
  cdae::DevBuf<float> acc3;  // H | HG | GU
  cdae::DevBuf<float> zd;    // Z | D
  int64_t scratch_users = 0;

  cdae::DevBuf<double> stage_d;
  cdae::DevBuf<float> stage_f;

  cdae::StatsDev* stats_d = nullptr;
  cdae::StatsDev* stats_h = nullptr;  // pinned
  int64_t launches = 0, h2d = 0, d2h = 0;

  // top-N table (cdae_topn_build)
  cdae::DevBuf<int32_t> topn_ids;     // [U][topk]
  cdae::DevBuf<float> topn_scores;    // [U][topk]
  cdae::DevBuf<float> topn_z;         // [U][ld] uncorrupted hidden vectors
  cdae::DevBuf<int> cand_id, cand_cnt, flag_d;
  cdae::DevBuf<float> cand_s;
  // tensor-core candidate path (topn_tc.cuh)
  cdae::DevBuf<uint16_t> tc_zb, tc_wb;   // bf16 operands [rows_pad][Kp]
  cdae::DevBuf<float> tc_wmax, tc_eps, tc_thr;
  cdae::DevBuf<int32_t> tc_redo;         // [unproven after sweep 1 | after sweep 2 | counter]
  cdae::DevBuf<float> tc_redo_thr;       // start threshold of each user of sweep 2
  const int32_t* tc_exact_list = nullptr;  // users left for the exact kernel (inside tc_redo)
  int64_t topn_pass2_users = 0;          // users that needed the second tensor sweep
  // probe pass (start thresholds of sweep 1 from the M items with the largest mean-user score)
  cdae::DevBuf<uint16_t> tc_probe_wb;    // [M][Kp] packed rows of the probe items
  cdae::DevBuf<float> tc_probe_keys;     // [2 I] sort keys in / out
  cdae::DevBuf<int32_t> tc_probe_ids;    // [2 I] item ids in / out (sorted by key, descending)
  cdae::DevBuf<int32_t> tc_probe_pos;    // [I] row of an item in the probe table, -1 = not in it
  cdae::DevBuf<uint32_t> tc_probe_bits;  // [users_pad][M / 32] rated bitmap over the probe table
  cdae::DevBuf<float> tc_probe_thr;      // [users] start thresholds
  cdae::DevBuf<float> tc_probe_zsum;     // [ld] column sums of the hidden vectors
  cdae::DevBuf<unsigned char> tc_probe_tmp;  // radix-sort scratch
  int topn_probe_items = 0;              // last cdae_topn_build: size of the probe table (0 = no probe pass)
  int64_t topn_tc_users = 0, topn_redo_users = 0;  // last cdae_topn_build: verified on the tensor path / redone exactly
  int topn_path = 0;                     // 0 fp32 CUDA cores, 1 tcgen05
  // full-item-decode training (fulldec_tc.cuh): bf16 operands and the loss-gradient matrix
  cdae::DevBuf<uint16_t> fd_zb, fd_wb, fd_g;   // [B_pad][Kp], [I_pad][Kp], [B_pad][I_pad]
  cdae::DevBuf<uint32_t> fd_bits;              // [B_pad][I_pad / 32] target bitmap of the slice
  cdae::DevBuf<float> fd_bias;                 // [I_pad] b' zero-padded (bias-outside mode, K + 2 > Kp)
  cdae::DevBuf<int64_t> test_rp_d;
  cdae::DevBuf<int32_t> test_col_d;
  std::vector<int32_t> topn_ids_h;    // host mirror for thread-safe lookups
  std::vector<float> topn_scores_h;
  int topn_k = 0;

  // cdae_profile: event pairs per launch, grouped by kernel class
  bool profiling = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_cls;
  size_t prof_used = 0;
  double prof_ms[CDAE_K_COUNT] = {0};
  int64_t prof_n[CDAE_K_COUNT] = {0};
  void prof_begin(int cls) {
    if (prof_used + 2 > prof_ev.size()) {
      for (int i = 0; i < 64; ++i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        prof_ev.push_back(e);
      }
      prof_cls.resize(prof_ev.size() / 2);
    }
    prof_cls[prof_used / 2] = cls;
    cudaEventRecord(prof_ev[prof_used], stream);
  }
  void prof_end() {
    cudaEventRecord(prof_ev[prof_used + 1], stream);
    prof_used += 2;
  }
  void prof_collect() {  // after a stream synchronise
    for (size_t i = 0; i + 1 < prof_used + 1 && i + 1 < prof_ev.size() && i < prof_used; i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, prof_ev[i], prof_ev[i + 1]) == cudaSuccess) {
        prof_ms[prof_cls[i / 2]] += ms;
        prof_n[prof_cls[i / 2]] += 1;
      }
    }
    prof_used = 0;
  }
};
