#!/usr/bin/env python3
"""Summarise one `ncu --set full` capture: writes profiles/<name>.csv (selected raw metrics) and prints the JSON entry that
profiles/ncu_summary.json keeps per workload (DRAM bytes per launch, executed FP64 flops as a fraction of the DFMA peak).

    python tools/ncu_extract.py gpurun_out/X.ncu-rep profiles/r02_ncu_full_X.csv [evals_per_launch] [kernel id]
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(Kernel Name|Block Size|Grid Size)$|pipe_tensor|dram__bytes_(read|write)\.sum|gpu__dram_throughput.avg.pct|gpu__time_duration.sum|"
    r"l1tex__data_pipe_lsu_wavefronts|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|launch__(block_size|grid_size|occupancy_limit|registers_per_thread|shared_mem_per_block_dynamic)|"
    r"sm__inst_executed_pipe_(fp64|lsu|alu|fma)\.avg\.pct|sm__pipe_fp64_cycles_active.avg.pct|sm__issue_active.avg.pct|sm__warps_active.avg.pct|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|smsp__inst_executed.sum$|sm__cycles_elapsed.avg$|sm__cycles_active.avg$|"
    r"smsp__sass_thread_inst_executed_op_(dfma|dadd|dmul)_pred_on.sum.per_cycle_elapsed$|sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained|"
    r"sass__inst_executed_local_(loads|stores)|smsp__inst_executed_op_local|l1tex__t_sectors_pipe_lsu_mem_local_op_(ld|st).sum$|"
    r"smsp__inst_executed_op_shfl|sm__throughput.avg.pct|smsp__cycles_active.avg$")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    evals = int(sys.argv[3]) if len(sys.argv) > 3 else None
    kid = int(sys.argv[4]) if len(sys.argv) > 4 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for h in hdr:
            if KEEP.search(h):
                w.writerow([h, m[h][0], m[h][1]])

    def num(k):
        return float(m[k][1].replace(",", "")) if k in m and m[k][1] not in ("", "n/a") else None

    def scaled(k):       # value in base units
        u, v = m[k]
        v = float(v.replace(",", ""))
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    dfma, dadd, dmul = (num(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed") for o in ("dfma", "dadd", "dmul"))
    entry = {"kernel_name": m["Kernel Name"][1], "kernel": kid, "evals_per_launch": evals,
             "dram_bytes_per_launch": scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum"),
             "duration_ms": num("gpu__time_duration.sum"), "registers": num("launch__registers_per_thread"),
             "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
             "issue_active_pct": num("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
             "lsu_data_pipe_pct": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
             "source": out}
    if None not in (dfma, dadd, dmul):
        # executed FP64 flops per elapsed cycle (2 per DFMA, 1 per DADD / DMUL thread instruction) over the DFMA peak of the
        # chip (64 lanes x 2 flop per SM and cycle, 148 SMs = 18944 flop/cycle)
        entry["executed_fp64_flop_frac"] = (2 * dfma + dadd + dmul) / (148 * 128)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
