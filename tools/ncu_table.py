"""One markdown table row per `ncu --set full` capture (CSV from `ncu -i X.ncu-rep --page raw --csv`): duration, clock,
FLOP-derived tensor fraction at the captured clock, the tensor-pipe counters, DRAM bytes.

    python tools/ncu_table.py name=flops:csv [name=flops:csv ...]
"""
import csv
import sys


def main():
    print("| capture | kernel | µs | SM GHz | FLOP/s ÷ dense peak @clock | `sm__pipe_tensor_cycles_active` % of elapsed | `…_realtime` (TPC triage) % | UTCHMMA issued (sum / min / max per SM) | DRAM read / write MB | DRAM % | regs | grid |")
    print("|---|---|---:|---:|---:|---:|---:|---|---|---:|---:|---:|")
    for arg in sys.argv[1:]:
        name, rest = arg.split("=", 1)
        flops, path = rest.split(":", 1)
        rows = list(csv.reader(open(path)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: vals[i] for i, h in enumerate(hdr)}
        u = {h: units[i] for i, h in enumerate(hdr)}
        g = lambda k: float(d[k].replace(",", ""))
        us = g("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u["gpu__time_duration.sum"], 1.0)
        ghz = g("sm__cycles_elapsed.avg.per_second") * {"Ghz": 1.0, "Mhz": 1e-3}.get(u["sm__cycles_elapsed.avg.per_second"], 1.0)
        mb = lambda k: g(k) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u[k], 1.0)
        peak = 148 * 8192 * ghz * 1e9          # dense bf16 FLOP/s at the captured clock: 4096 MAC/clk/SM
        frac = float(flops) / (us * 1e-6) / peak
        kern = d["Kernel Name"].split("(")[0].replace("void <unnamed>::", "")
        print(f"| {name} | `{kern}` | {us:.1f} | {ghz:.3f} | {frac:.3f} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g('sm__inst_executed_pipe_tensor.sum'):.0f} / {g('sm__inst_executed_pipe_tensor.min'):.0f} / {g('sm__inst_executed_pipe_tensor.max'):.0f} | "
              f"{mb('dram__bytes_read.sum'):.0f} / {mb('dram__bytes_write.sum'):.0f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{d['launch__registers_per_thread']} | {d['launch__grid_size']} |")


main()
