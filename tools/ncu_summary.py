"""Condenses an .ncu-rep (ncu --set full) into the per-kernel numbers the roofline claims use.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
    # L2 side of the gather / reduction kernels
    ("lts__t_sectors.sum", "l2_sectors"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_sectors_read"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "l2_sectors_write"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "l2_sectors_red"),
    ("lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed", "l2_red_pct"),
    ("lts__t_sectors_srcunit_tex_op_atom.sum", "l2_sectors_atom"),
    ("lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "l2_atomic_unit_pct"),
    ("lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "l2_sectors_pct"),
    ("lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed", "l2_to_xbar_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "xbar2l1_read"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum", "l12xbar_write"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "red_requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "red_sectors"),
    ("sm__cycles_elapsed.max", "cycles"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    tensor_cols = [h for h in hdr if "pipe_tensor" in h and "pct_of_peak_sustained_active" in h]
    print("# %s" % path)
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print(name)
        for key, label in WANT:
            if key in idx and r[idx[key]] not in ("", "n/a"):
                print("    %-16s %s %s" % (label, r[idx[key]], units[idx[key]]))
        for h in tensor_cols:
            if r[idx[h]] not in ("", "n/a", "0"):
                print("    %-16s %s %s" % (h[:60], r[idx[h]], units[idx[h]]))
        # execution pipes above 2 % (which pipe an issue-bound kernel is bound by)
        pipes = []
        for h in hdr:
            if (h.startswith("sm__inst_executed_pipe_") or (h.startswith("sm__pipe_") and "cycles_active" in h)) and \
                    h.endswith(".avg.pct_of_peak_sustained_active") and "tensor" not in h and r[idx[h]] not in ("", "n/a"):
                try:
                    pipes.append((float(r[idx[h]].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(pipes, reverse=True)[:8]:
            if v >= 2.0:
                print("    pipe  %-44s %.1f %%" % (h.replace(".avg.pct_of_peak_sustained_active", ""), v))
        # warp-stall breakdown: cycles a warp waits per issued instruction, by reason (top 6)
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[idx[h]] not in ("", "n/a"):
                try:
                    stalls.append((float(r[idx[h]].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:6]:
            print("    stall %-28s %.2f warps per issued instruction" % (h.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    main(sys.argv[1])
