"""Summarises an .ncu-rep (read on the CPU box with `ncu -i`) into a small markdown table for profiles/."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
    # the L1TEX data pipe (shared-memory and global wavefronts): the busiest unit of pass B
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed",
    "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg",
    "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of `{rep.split('/')[-1]}` (per launch; cold-cache, serialised)", ""]
    for r in rows[2:]:
        lines.append(f"## {r[hdr.index('Kernel Name')]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"| {w} | {r[i]} | {units[i]} |")
        lines.append("")
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main(*sys.argv[1:])
