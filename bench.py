#!/usr/bin/env python
"""users/sec of CDAE training (BASELINE.json metric) on N B200s, plus the CPU reference arm.

A step = ONE EPOCH (CDAE::train_one_iteration, cdae.hpp:136-146) over a synthetic data set of one of
BASELINE.json's configurations (weak scaling: the per-GPU user count is fixed):

  B  (default, the configuration the metric is quoted on)  100,000 users per GPU x 50,000 items,
     ~30 train items per user, K=50, num_neg=5, tied weights, sampled decode
  C  138,000 users x 27,000 items, ~145 items per user, K=200, FULL-item decode (tcgen05), asymmetric
  D  1M x 200K over 8 GPUs = 125,000 users per GPU x 200,000 items, K=100, num_neg=5, tied
  E  500K x 100K over 8 GPUs = 62,500 users per GPU x 100,000 items, K=256, full-item decode

all with q=0.5 scaled, CROSS_ENTROPY, lambda=.01, lr=.1, AdaGrad beta=1, user factor on.  `value`
times the epoch with the CSR resident in HBM (CUDA events on the engine's stream, no per-kernel
events in that region); `e2e` times cdae_train_epoch_csr, which takes the CSR from pinned HOST
memory on every call and reads the epoch statistics back; a third, profiled pass over the same
inputs gives the per-kernel times the roofline is computed from.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config B|C|D|E]      (torchrun for N > 1)
  python bench.py --impl reference ...   the reference's own CPU path (oracle/_ref)

Without --config the line is config B and carries the other configurations as `configs`: C (full
138K users) at N = 1; D and E (their per-GPU shards) at N = 8 (or with --extra D,E at any N).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

UNIT = "users/s"
SEED = 20141119

CONFIGS = {
    "B": dict(users_per_gpu=100_000, items=50_000, mean=30.0, K=50, num_neg=5, asym=False, full=False,
              batch=16384, gpus=1, metric="users/sec CDAE training (K=50, Yelp-scale)",
              what="synthetic %dK users x 50K items, K=50, neg-sample=5, %dxB200"),
    "C": dict(users_per_gpu=138_000, items=27_000, mean=145.0, K=200, num_neg=5, asym=True, full=True,
              batch=0, gpus=1, metric="users/sec CDAE training (K=200, MovieLens-20M-scale, full-item decode)",
              what="MovieLens-20M-scale synthetic (%dK x 27K), K=200, full-item decode, %dxB200"),
    "D": dict(users_per_gpu=125_000, items=200_000, mean=30.0, K=100, num_neg=5, asym=False, full=False,
              batch=16384, gpus=8, metric="users/sec CDAE training (K=100, Yelp-scale 1M x 200K)",
              what="Yelp-scale synthetic %dK x 200K, K=100, neg-sample=5, user-sharded %dxB200"),
    "E": dict(users_per_gpu=62_500, items=100_000, mean=50.0, K=256, num_neg=5, asym=True, full=True,
              batch=0, gpus=8, metric="users/sec CDAE training (K=256, 500K x 100K, full-item decode bf16)",
              what="full-item decode %dK x 100K, K=256, bf16 tcgen05 path, %dxB200"),
}


def model_cfg(c):
    return dict(lambda_=0.01, learn_rate=0.1, corruption_ratio=0.5, beta=1.0, loss="CE",
                num_dim=c["K"], num_neg=c["num_neg"], num_corruptions=1, using_adagrad=True,
                asymmetric=c["asym"], user_factor=True, linear=False, scaled=True,
                linear_function=False, tanh=False)


def config_dict(name, world, batch_users_global, sms=148):
    """The `config` object of the JSON line — the same function serves both arms, so the reference
    arm and ours describe one workload."""
    c = CONFIGS[name]
    U = c["users_per_gpu"] * world
    return {"workload": c["what"] % (U // 1000, world), "name": name, "users": U, "items": c["items"],
            "mean_train_items_per_user": c["mean"], "num_dim": c["K"],
            "decode": "all items" if c["full"] else "positives + %d sampled negatives each" % c["num_neg"],
            "weights": "asymmetric" if c["asym"] else "tied", "loss": "CE", "optimizer": "AdaGrad beta=1, lr=0.1, lambda=0.01",
            "corruption": "q=0.5, scaled", "step": "one epoch = CDAE::train_one_iteration over all users",
            "seed": SEED}


def peaks():
    """Driver-measured peaks (MEASURED_PEAKS.json), else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def decode_kernel_name(K):
    """The <G,NV> instantiation DISPATCH_LD (csrc/api.cu) picks for this K."""
    lines = (K + 31) // 32
    lines = lines if lines <= 4 else 6 if lines <= 6 else 8 if lines <= 8 else 12 if lines <= 12 else 16
    g, nv = ((8, lines) if lines <= 4 else (16, 3) if lines <= 6 else (16, 4) if lines <= 8
             else (32, 3) if lines <= 12 else (32, 4))
    return "decode_kernel<%d,%d,TRAIN,SAMPLED,CE>" % (g, nv), lines * 32


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get(kernel)
    return (d["dram_bytes_per_launch"], d["source"]) if d else (None, None)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# data
def rank_dataset(name, rank, world, batch_global):
    """This rank's view of the configuration's data set.  One GPU: the whole set.  A process group:
    only the rows of the users the rank trains are generated (an independent block per rank over one
    shared item popularity model); the other users' rows are empty in its CSR."""
    from cdae_b200 import synth
    c = CONFIGS[name]
    U = c["users_per_gpu"] * world
    if world == 1:
        if name == "B":
            return synth.make_dataset(U, c["items"], c["mean"], seed=SEED)       # the r01 arrays
        return synth.make_blocked_dataset(U, c["items"], c["mean"], seed=SEED)
    return synth.make_sharded_dataset(U, c["items"], c["mean"], batch_global, rank, world, seed=SEED)


# --------------------------------------------------------------------------------------------
# the CPU arm
def cpu_reference_rate(name, data, n_sample, steps=1, warmup=0, quiet=True):
    """users/s of the reference's own (single-threaded, fp64) train loop on the first n_sample users.
    Sampled decode (B, D): the VERBATIM reference headers (oracle/_ref; the default build and a
    -march=x86-64-v3 build are both timed, the faster one is reported), else the C port.  Full-item
    decode (C, E) has no reference function: the oracle's full mode is timed.
    Returns (value, info, times)."""
    from oracle import oracle as orc
    c = CONFIGS[name]
    mcfg = model_cfg(c)
    rp, col = data["train_row_ptr"], data["train_col"]
    n = int(min(n_sample, data["U"]))
    if c["full"]:
        o = orc.Oracle(mcfg, data["U"], data["I"], rp, col)
        o.init_params(SEED)
        times = []
        for s in range(warmup + steps):
            t = time.perf_counter()
            o.train_epoch_full(SEED, s, n, 0, 0, n)
            times.append(time.perf_counter() - t)
        times = times[warmup:]
        info = dict(kind="port", cores=1,
                    sample="C port of the frozen-batch step with all items as outputs (oracle/; the reference has no "
                           "full-item-decode training), first %d of %d users as one minibatch, fp64, 1 thread" % (n, data["U"]))
        return n * len(times) / sum(times), info, times
    sub_rp = np.ascontiguousarray(rp[:n + 1])
    sub_col = np.ascontiguousarray(col[:rp[n]])
    if orc.have_reference():
        col2, i_seen = orc.first_seen_relabel(sub_rp, sub_col)
        builds = [(so, label) for so, label in ((orc.REF_SO, "-O3"), (orc.REF_V3_SO, "-O3 -march=x86-64-v3")) if os.path.exists(so)]
        if len(builds) > 1:                                  # pick the faster build on a 1,000-user prefix
            k = min(n, 1000)
            k_rp, k_col = np.ascontiguousarray(rp[:k + 1]), np.ascontiguousarray(col[:rp[k]])
            k_col2, k_seen = orc.first_seen_relabel(k_rp, k_col)
            trial = []
            for so, label in builds:
                ref = orc.Reference(mcfg, k, k_seen, k_rp, k_col2, quiet=quiet, so=so)
                orc.Reference.seed(SEED, SEED, so=so)
                trial.append(min(ref.train_user_range(0, k) for _ in range(2)))
                del ref
            builds = [builds[int(np.argmin(trial))]]
        so, label = builds[0]
        ref = orc.Reference(mcfg, n, i_seen, sub_rp, col2, quiet=quiet, so=so)
        orc.Reference.seed(SEED, SEED, so=so)
        times = [ref.train_user_range(0, n) for _ in range(warmup + steps)][warmup:]
        rate = n * len(times) / sum(times)
        info = dict(kind="reference", cores=1, build="g++ " + label + " (the faster of the builds present)",
                    sample="verbatim reference headers (oracle/_ref; Eigen/Boost/glog are this repo's stand-ins: "
                           "plain loops, real Eigen may differ), first %d of %d users = %d of %d items seen (ids relabelled), "
                           "fp64, 1 thread (the reference trains single-threaded, cdae.hpp:136-146)"
                           % (n, data["U"], i_seen, data["I"]))
        return rate, info, times
    o = orc.Oracle(mcfg, data["U"], data["I"], rp, col)
    o.init_params(SEED)
    times = []
    for s in range(warmup + steps):
        t = time.perf_counter()
        o.train_epoch(SEED, s, 1, 0, n)
        times.append(time.perf_counter() - t)
    times = times[warmup:]
    info = dict(kind="port", cores=1, sample="C port of the reference loop (oracle/), first %d of %d users, 1 thread" % (n, data["U"]))
    return n * len(times) / sum(times), info, times


def cpu_multicore_rate(name, data, n_sample):
    """BASELINE.md §3 baseline (ii): OUR Hogwild port of the reference loop (oracle/, fp64, lock-free
    user-parallel over all host cores) — labelled as a port, the reference itself has no parallel trainer."""
    from oracle import oracle as orc
    c = CONFIGS[name]
    if c["full"]:
        return None
    cores = os.cpu_count() or 1
    n = int(min(n_sample * min(cores, 16), data["U"]))
    o = orc.Oracle(model_cfg(c), data["U"], data["I"], data["train_row_ptr"], data["train_col"])
    o.init_params(SEED)
    o.train_epoch_hogwild(SEED, 0, cores, 0, min(n, 2000))                      # thread start-up
    t = time.perf_counter()
    o.train_epoch_hogwild(SEED, 1, cores, 0, n)
    dt = time.perf_counter() - t
    return dict(value=n / dt, unit=UNIT, cores=cores, kind="port",
                sample="oracle/ Hogwild (lock-free user-parallel) port of the reference loop, fp64, first %d of %d users, %d threads"
                       % (n, data["U"], cores))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    orc.build(ref=False)
    name = args.config or "B"
    c = CONFIGS[name]
    data = rank_dataset(name, 0, 1, 0)
    # calibrate so that the whole (warmup + steps) run stays within ~2 minutes
    rate, _, _ = cpu_reference_rate(name, data, 64 if c["full"] else 1000)
    total = max(1, args.steps + args.warmup)
    # (the 1,000-user calibration runs cache-resident and over-estimates the rate of a longer prefix by ~2x)
    n = int(max(16 if c["full"] else 500, min(c["users_per_gpu"], args.cpu_sample, rate * 40.0 / total)))
    value, info, times = cpu_reference_rate(name, data, n, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    info["value"] = value
    info["unit"] = UNIT
    info["sample_users_per_step"] = n
    line = {"impl": "reference", "metric": c["metric"], "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(name, 1, 0),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------
# our arm
class Ctx:
    pass


def measure(name, ctx, args, primary):
    """One configuration on this process group: returns (result dict for rank 0, data set)."""
    import torch
    import torch.distributed as dist
    from cdae_b200 import CDAE, CDAEConfig
    c = CONFIGS[name]
    world, rank, local = ctx.world, ctx.rank, ctx.local
    sms = ctx.sms
    batch_per_gpu = c["batch"] if c["batch"] > 0 else 128 * sms
    if name == "B" and args.batch_users:
        batch_per_gpu = args.batch_users
    batch_global = batch_per_gpu * world                        # weak scaling: the global minibatch grows with N
    U, I, K = c["users_per_gpu"] * world, c["items"], c["K"]
    mcfg = model_cfg(c)
    data = rank_dataset(name, rank, world, batch_global)
    rp, col = data["train_row_ptr"], data["train_col"]
    m = CDAE(CDAEConfig(batch_users=batch_global, device=local, full_decode=c["full"], **mcfg)).reset(U, I, rp, col)
    allreduce = "none"
    if world > 1:
        uid = [CDAE.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        m.dist_init(rank, world, uid[0])
        allreduce = "NCCL all-reduce + replicated apply"
        if ctx.p2p:
            def gather(b):
                box = [None] * world
                dist.all_gather_object(box, b)
                return box
            # NVLS wins where the peer-memory kernel is NVLink-bound (8 GPUs: 71 vs 85 us per 13.6 MB step); on 2-4 GPUs the
            # peer kernel is faster (52 vs 73 us, 67 vs 69 us) — profiles/r02_c_combine_probe_*.json
            want_nvls = args.allreduce == "nvls" or (args.allreduce == "auto" and world >= 8)
            if want_nvls and os.environ.get("CDAE_B200_NVLS", "1") != "0" and m.dist_mc_init(rank, world, gather):
                allreduce = ctx.nvls_label
            else:
                m.dist_p2p_init(gather)
                allreduce = ctx.p2p_label
    m.init_params(SEED)
    rp_pin, col_pin = m.pinned_array(rp), m.pinned_array(col)
    steps = args.steps if primary else max(3, min(args.steps, 5))
    warmup = args.warmup if primary else max(3, min(args.warmup, 3))

    def barrier():
        torch.cuda.synchronize()
        m.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_epoch(epoch, host_csr):
        ctx.flush.zero_()                   # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        t = time.perf_counter()
        st = m.train_one_iteration(seed=SEED, epoch=epoch, csr=(rp_pin, col_pin) if host_csr else None)
        return st, time.perf_counter() - t

    epoch = 0
    for _ in range(warmup):
        one_epoch(epoch, False)
        epoch += 1
    # ---- timed region 1: CSR resident in HBM, device time (CUDA events on the engine stream)
    sampler = ClockSampler(local) if (rank == 0 and primary) else None
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, outputs, users, loss = 0.0, 0, 0, 0, 0.0
    for _ in range(steps):
        st, _w = one_epoch(epoch, False)
        epoch += 1
        dev_ms += st.device_ms
        launches += st.kernel_launches
        outputs += st.outputs
        users += st.user_steps
        loss = st.loss_sum
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None
    # ---- timed region 2: end to end through the host-CSR call (wall clock incl. H2D + D2H)
    for _ in range(2):
        one_epoch(epoch, True)
        epoch += 1
    barrier()
    e2e_s, h2d, d2h = 0.0, 0, 0
    for _ in range(steps):
        st, w = one_epoch(epoch, True)
        epoch += 1
        e2e_s += w
        h2d, d2h = st.h2d_bytes, st.d2h_bytes
    barrier()
    # ---- region 3: the same epochs with an event pair around every kernel launch (per-kernel times)
    m.profile(True)
    prof_ms, prof_out = 0.0, 0
    for _ in range(steps):
        st, _w = one_epoch(epoch, False)
        epoch += 1
        prof_ms += st.device_ms
        prof_out += st.outputs
    barrier()
    prof = m.profile_get()
    m.profile(False)

    tv = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([users, outputs, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)       # max over ranks
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = tv.tolist()
    users_all, outputs_all, launches_all = cnt.tolist()
    res = None
    if rank == 0:
        pk = peaks()
        value = users_all / (dev_ms / 1e3)
        share = {k: v[0] / prof_ms for k, v in prof.items() if v[1]}
        res = {"metric": c["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "bf16 operands, f32 accumulate (f32 everywhere outside the decode)" if c["full"] else "f32",
               "data": "synthetic"}
        cfgd = config_dict(name, world, batch_global, sms)
        res["config"] = cfgd
        res["run"] = {"train_nnz_rank0": int(len(col)), "batch_users_global": batch_global, "batch_users_per_gpu": batch_per_gpu,
                      "l2": "256 MB flush write between timed iterations",
                      "parallelism": "dp%d: users sharded, item-side gradients combined once per minibatch (%s)" % (world, allreduce),
                      "loss_last_epoch_rank0": loss}
        res["e2e"] = {"value": users_all / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / steps}
        res["gpu_launches"] = int(launches_all)
        if clocks is not None:
            res["clocks"] = clocks
        res["kernel_ms_share"] = share
        res["kernel_share_note"] = ("from a separate profiled pass (event pair around every launch, %.3f ms per step against "
                                    "%.3f unprofiled)" % (prof_ms / steps, dev_ms / steps))
        if c["full"]:
            res["roofline"] = tensor_roofline(c, prof, steps, users_all / world / steps, dev_ms / steps, pk)
        else:
            res["roofline"] = decode_roofline(m, c, prof, prof_out, pk)
    ctx.last_model = m
    return res, data


def decode_roofline(m, c, prof, outputs_profiled, pk):
    """The dominant kernel of the sampled path.  §8(d) algorithmic bytes: outputs * (P*4K + P*4 + 4),
    P = 4 row passes with AdaGrad.  When the item tables + gradients fit the L2 (config B: 38 MB of 126)
    the kernel does not move those bytes through HBM, so the bound is the L2: `peak` is then measured
    live by cdae_probe_l2 (the same row loads + vector reductions, nothing else) and `achieved` counts
    the bytes the kernel really requests from L2 (padded rows: one read + one reduction per output)."""
    K, I = c["K"], c["items"]
    kname, ld = decode_kernel_name(K)
    dec_ms, dec_n = prof["decode"]
    if not dec_n:
        return None
    out_per_launch = outputs_profiled / dec_n
    avg_ms = dec_ms / dec_n
    alg = out_per_launch * (4 * 4 * K + 4 * 4 + 4)
    l2_bytes = out_per_launch * (2 * ld * 4 + 12)               # row read + row reduction (+ b' read, gb' reduction, item id)
    working_set = I * ld * 4 * 3
    visits = int(out_per_launch)
    both, both_ms = m.probe_l2(I, 3, visits)
    rd, _ = m.probe_l2(I, 1, visits)
    red, _ = m.probe_l2(I, 2, visits)
    traffic, traffic_src = ncu_traffic("decode_kernel" if c["items"] == 50_000 else "decode_kernel_D")
    l2_resident = working_set < 100e6
    ach_alg = alg / (avg_ms / 1e3) / 1e9
    ach_l2 = l2_bytes / (avg_ms / 1e3) / 1e9
    r = {"kernel": kname, "launches": dec_n, "avg_launch_ms": avg_ms, "unit": "GB/s",
         "algorithmic_bytes_per_launch": alg, "achieved_algorithmic": ach_alg,
         "hbm_peak": pk["hbm"], "peak_source": pk["source"],
         "l2_bytes_per_launch": l2_bytes, "achieved_l2": ach_l2,
         "peak_l2": both, "peak_l2_read_only": rd, "peak_l2_reduction_only": red,
         "peak_l2_source": "cdae_probe_l2 run by this process: uniformly random rows of a %d x %d fp32 table, ld.v4 + red.v4 per row, "
                           "same lane geometry, %d row visits per launch (%.3f ms)" % (I, ld, visits, both_ms),
         "working_set_bytes": working_set, "traffic": traffic, "traffic_source": traffic_src}
    # The governing roofline of this kernel is the memory system's load + vector-REDUCTION throughput on a table of
    # this size (reductions resolve in L2 at ~5.4 TB/s chip-wide whether they are issued as red.v4 or as bulk
    # cp.reduce, profiles/r02_b_l2_probe.json), measured by the probe; HBM figures ride along.
    r.update({"bound": "l2", "achieved": ach_l2, "peak": both, "frac": ach_l2 / both,
              "hbm_frac_of_algorithmic_bytes": ach_alg / pk["hbm"]})
    if l2_resident:
        r["note"] = ("W + its AdaGrad state + the gradient buffer (%.0f MB) are L2-resident, so the §8(d) algorithmic bytes never "
                     "reach HBM (hbm_frac_of_algorithmic_bytes is NOT a roofline fraction); the bound is the L2's load + "
                     "vector-reduction throughput, measured by the probe" % (working_set / 1e6))
    else:
        r["note"] = ("working set %.0f MB exceeds the 126 MB L2: the probe runs on the same table size, so its mix of L2 hits and HBM "
                     "misses is the kernel's; §8(d)'s P = 4 row passes charge this kernel two passes that apply / the fused combine "
                     "step make, which is why hbm_frac_of_algorithmic_bytes can exceed 1" % (working_set / 1e6))
    return r


def tensor_roofline(c, prof, steps, users_rank, step_ms, pk):
    """Full-item decode: the three tcgen05 contractions, 6*I*K flops per user, against the
    SUSTAINED bf16 peak (the kernels run back to back inside a long step).  users_rank = users one rank
    trains per epoch."""
    per = {k: v[0] / steps for k, v in prof.items() if v[1]}
    tens = [k for k in ("fd_score", "fd_hidden", "fd_itemgrad") if k in per]
    tens_ms = sum(per[k] for k in tens)
    flops = 6.0 * users_rank * c["items"] * c["K"]
    ach = flops / (tens_ms / 1e3) / 1e12
    return {"bound": "tensor", "kernel": " + ".join(tens), "achieved": ach, "peak": pk["tf_sustained"],
            "peak_source": pk["source"] + ", sustained", "peak_burst": pk["tf_burst"], "unit": "TFLOP/s",
            "frac": ach / pk["tf_sustained"], "flops": "6*U*I*K per epoch (SURVEY 8d), per GPU",
            "per_kernel_ms_per_epoch": per,
            "per_kernel_tflops": {k: (flops / 3.0) / (per[k] / 1e3) / 1e12 for k in tens},
            "whole_step_tflops": flops / (step_ms / 1e3) / 1e12,
            "whole_step_frac": flops / (step_ms / 1e3) / 1e12 / pk["tf_sustained"],
            "traffic": ncu_traffic("fd_gemm_kernel")[0]}


def topn_section(m, c, U, pk, sm_mhz=None, sms=148):
    """Secondary measurement (not part of the timed training step): CDAE::recommend for all users,
    i.e. the full-item decode on tcgen05 (csrc/topn_tc.cuh), against the measured bf16 peak."""
    K, ITEMS = c["K"], c["items"]
    m.pre_recommend(10)                                       # warm-up (allocations, tensor maps)
    m.profile(True)
    reps = 3
    for _ in range(reps):
        m.pre_recommend(10)
    prof = m.profile_get()
    m.profile(False)
    path, verified, redone = m.topn_stats()
    ms = prof["topn"][0] / reps
    kp = (K + 2 + 63) // 64 * 64
    alg = 2.0 * U * ITEMS * K                                  # 2*I*K flops per user
    return {"what": "CDAE::recommend, all users x all items, top-10 (cdae_topn_build)",
            "path": "tcgen05 bf16 + exact fp64 re-rank" if path == 1 else "fp32 CUDA cores",
            "users_per_s": U / (sum(prof[k][0] for k in ("gather", "activate", "topn", "topn_pack", "topn_rerank", "topn_exact")) / reps / 1e3),
            "candidate_kernel_ms": ms, "verified_users": verified, "redone_exact_users": redone,
            "probe_items": m.topn_probe_items(),               # > 0: sweep 1 started from probe thresholds (both launches are in candidate_kernel_ms)
            # At small K the sweep is not tensor-bound: every score is read out of TMEM once, and the guide's measured
            # TMEM read rate (B300_MICROARCH.md "LDTM throughput": 64 B per clock per SM, measured on sm_103 — not
            # re-measured on this B200) is the floor: U * I * 4 bytes against SMs * 64 B * SM clock.
            "roofline_epilogue": (lambda clk: {"bound": "tmem_read", "unit": "GB/s",
                                               "achieved": U * ITEMS * 4.0 / (ms / 1e3) / 1e9 if ms > 0 else None,
                                               "peak": sms * 64.0 * clk * 1e6 / 1e9,
                                               "peak_source": "B300_MICROARCH.md LDTM 64 B/clk/SM x %d SMs x %.0f MHz (guide constant, not re-measured)" % (sms, clk),
                                               "frac": (U * ITEMS * 4.0 / (ms / 1e3) / 1e9) / (sms * 64.0 * clk * 1e6 / 1e9) if ms > 0 else None})(
                float(sm_mhz or 1965.0)),
            "roofline": {"bound": "tensor", "kernel": "topn_tc_kernel<%d>" % (kp // 64),
                         "achieved": alg / (ms / 1e3) / 1e12 if ms > 0 else None, "peak": pk["tf_burst"],
                         "peak_source": pk["source"] + ", burst", "unit": "TFLOP/s",
                         "frac": alg / (ms / 1e3) / 1e12 / pk["tf_burst"] if ms > 0 else None,
                         "executed_tflops_padded_k": 2.0 * U * ITEMS * kp / (ms / 1e3) / 1e12 if ms > 0 else None,
                         "scores_per_s": U * ITEMS / (ms / 1e3) if ms > 0 else None,
                         "traffic": ncu_traffic("topn_tc_kernel")[0]}}


def run_ours(args):
    import torch
    import torch.distributed as dist

    ctx = Ctx()
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        sys.exit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx.sms = torch.cuda.get_device_properties(local).multi_processor_count
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    # the combine step: one fused kernel over NVLink peer memory (default), or NCCL all-reduce + replicated apply
    p2p_env = os.environ.get("CDAE_B200_P2P", "")
    ctx.p2p = world > 1 and args.allreduce != "nccl" and p2p_env != "0"
    ctx.nvls_label = ("one kernel through the NVSwitch multicast engine (NVLS): multimem.ld_reduce of the rank's gradient slice, its slice "
                      "of the AdaGrad step, multimem.st of the updated parameters")
    ctx.p2p_label = ("one kernel over NVLink peer memory: reduce-scatter by peer loads, the rank's slice of the AdaGrad step, "
                     "all-gather of the updated parameters by peer stores")

    primary = args.config or "B"
    extras = []
    if args.extra:
        extras = [x for x in args.extra.split(",") if x and x != primary]
    elif not args.config and not args.no_extra:
        extras = ["C"] if world == 1 else ["D", "E"] if world == 8 else []
    line, data = measure(primary, ctx, args, True)
    m = ctx.last_model
    c = CONFIGS[primary]
    if rank == 0:
        if world == 1 and primary == "B" and not args.no_topn:
            line["topn"] = topn_section(m, c, c["users_per_gpu"], peaks(),
                                        sm_mhz=(line.get("clocks") or {}).get("sm_mhz"), sms=ctx.sms)
    m.close()
    ctx.last_model = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build(ref=False)
        v, info, _ = cpu_reference_rate(primary, data, args.cpu_sample if not c["full"] else 48)
        info["value"] = v
        info["unit"] = UNIT
        line["cpu_baseline"] = info
        mc = cpu_multicore_rate(primary, data, args.cpu_sample)
        if mc:
            line["cpu_baseline_multicore"] = mc
    del data
    for name in extras:
        res, _d = measure(name, ctx, args, False)
        ctx.last_model.close()
        ctx.last_model = None
        if rank == 0:
            line.setdefault("configs", {})[name] = res
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


class OneLineStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints
    "NCCL version ..." from C code at communicator init), so file descriptor 1 is pointed at stderr
    for the whole run and the JSON line goes to the saved descriptor."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())


OUT = None


def emit(obj):
    OUT.emit(json.dumps(obj))


def main():
    global OUT
    OUT = OneLineStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="BASELINE.json configuration (default B)")
    ap.add_argument("--extra", default=None, help="comma list of further configurations to add under `configs`")
    ap.add_argument("--no-extra", action="store_true", help="only the primary configuration")
    ap.add_argument("--batch-users", type=int, default=0, help="config B: users per minibatch per GPU (default 16384)")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nccl", "p2p", "nvls"],
                    help="combine step: auto = NVLS kernel where multicast is available, else the peer-memory kernel")
    ap.add_argument("--cpu-sample", type=int, default=30000,
                    help="users in the bounded CPU-baseline sample (about 10-15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-topn", action="store_true", help="skip the recommend (full-item decode) section")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
