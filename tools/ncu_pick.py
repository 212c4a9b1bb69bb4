"""Pick the roofline-relevant metrics out of `ncu -i X.ncu-rep --page raw --csv` (one row per profiled launch).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tools/ncu_pick.py > profiles/rNN_prof.md
"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__cluster",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]


def main():
    rows = list(csv.reader(ln for ln in sys.stdin if ln.startswith('"')))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kn = col.get("Kernel Name")
    for r in rows[2:]:
        print(f"### launch {r[col['ID']]}: `{r[kn][:70]}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        print()
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for h in hdr:
            if any(h.endswith(w) or h == w for w in WANT):
                print(f"| `{h}` | {r[col[h]]} | {units[col[h]]} |")
        print()


if __name__ == "__main__":
    main()
