"""Key metrics of every kernel in an ncu report (`--set full`), as markdown.  usage: python tools/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sectors_op_read.sum",
    "lts__t_sectors_op_write.sum", "lts__t_sectors_srcunit_tex.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "launch__local_mem_per_thread" ,
]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print("### `%s`\n" % d.get("Kernel Name", "?")[:140])
        print("| metric | value | unit |\n|---|---:|---|")
        for w in WANT:
            if w in d and d[w] != "":
                print("| %s | %s | %s |" % (w, d[w], u[w]))
        tens = [h for h in hdr if "tensor" in h and d.get(h) not in ("", "0", None) and h not in WANT]
        for h in tens[:8]:
            print("| %s | %s | %s |" % (h, d[h], u[h]))
        print()


if __name__ == "__main__":
    main()
