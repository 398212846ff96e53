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


if __name__ == "__main__":
    main(sys.argv[1])
