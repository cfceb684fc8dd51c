"""Summary of an `ncu --set full` report for profiles/: per captured launch the metrics the roofline discussion uses
(duration, grid, registers, shared memory, DRAM bytes, L2 hit rate, occupancy, issue slots, FP64 / DMMA pipes) and the
top warp-state samples.
    python tools/summarise_ncu.py gpurun_out/x.ncu-rep "header line" > profiles/x_summary.txt"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    print(sys.argv[2] if len(sys.argv) > 2 else rep)
    print("(B200, sm_100a, round 2)\n")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(hdr):
        col.setdefault(h.split(".TriageCompute.")[-1] if ".Triage" in h else h, i)
    for r in data:
        print("== " + r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in col and r[col[k]] != "":
                print("   %-82s %s %s" % (k, r[col[k]], units[col[k]]))
        # warp states: smsp__pcsamp_warps_issue_stalled_* sample counts
        st = {}
        for h, i in col.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i] not in ("", "0"):
                try:
                    st[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(st.values())
        if tot > 0:
            top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
            print("   warp-state samples: " + ", ".join("%s %d%%" % (k, round(100 * v / tot)) for k, v in top))


if __name__ == "__main__":
    main()
