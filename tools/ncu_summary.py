"""Key metrics of an `ncu --set full` report as a small markdown table (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep "title" >> profiles/...md"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active (% of peak, active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (% of peak, elapsed)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("### %s\n" % title)
    print("`%s`\n" % path.split("/")[-1])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("kernel: `%s`\n" % name[:110])
        print("| metric | value |\n|---|---|")
        for key, label in WANT:
            for i, h in enumerate(hdr):
                if h == key:
                    print("| %s (`%s`) | %s %s |" % (label, key, r[i], units[i]))
        print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
