"""Summarise an `ncu --page raw --csv` export: one block per launch with the metrics DESIGN.md / bench.py quote."""
import csv
import re
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram bytes read"),
    ("dram__bytes_write.sum", "dram bytes written"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (active)"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long scoreboard %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
]


def main(path, limit=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {h: i for i, h in enumerate(hdr)}
    n = 0
    for r in data:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        print("launch %s  %s  grid %s block %s" % (r[col["ID"]], name, r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        for key, label in WANT:
            if key in col:
                print("  %-28s %s %s" % (label, r[col[key]], units[col[key]]))
        n += 1
        if limit and n >= limit:
            break


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
