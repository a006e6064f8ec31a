"""Summarise an `ncu --set full` report of conv_tcgen05 launches: one line per launch with the metrics DESIGN.md / bench.py cite,
and (optionally) the DRAM traffic of the largest launch as JSON for bench.py's roofline.traffic.
usage: python tools/ncu_summary.py <report.ncu-rep> <summary.txt> [traffic.json [launch index [description]]]"""
import csv
import json
import subprocess
import sys

METRICS = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
]


def main(rep, out_txt, out_json=None, index=None, desc=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    with open(out_txt, "w") as f:
        f.write(f"# {rep}: ncu --set full --clock-control none, {len(rows) - 2} launches (cold cache, serialised)\n")
        f.write(" | ".join(m for m, _ in cols) + "\n")
        f.write(" | ".join(units[i] for _, i in cols) + "\n")
        for r in rows[2:]:
            f.write(" | ".join(r[i] for _, i in cols) + "\n")
    if out_json:
        def num(r, name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(units[i], 1.0)
        best = rows[2 + int(index)] if index is not None else max(rows[2:], key=lambda r: num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"))
        json.dump({"source": out_txt + " (ncu --set full --clock-control none; launch #%s of the captured step)" % (index if index is not None else "max-traffic"),
                   "kernel": best[hdr.index("Kernel Name")] + " " + (desc or "(launch with the largest DRAM traffic of the captured step)"),
                   "dram_bytes_read": int(num(best, "dram__bytes_read.sum")), "dram_bytes_write": int(num(best, "dram__bytes_write.sum")),
                   "duration_us": best[hdr.index("gpu__time_duration.sum")]}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:6])
