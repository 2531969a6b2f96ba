#!/usr/bin/env python3
"""profiles/r2_kernel_traffic.json from an `ncu --set full` capture of one C3 mat-vec (tools/gpu_final_n1.sh):
per kernel class the DRAM bytes of one launch (dram__bytes_read.sum + dram__bytes_write.sum) and the counters the
DESIGN.md kernel table quotes.  Usage: python tools/make_traffic_table.py gpurun_out/r2_final_full.ncu-rep"""
import csv
import json
import subprocess
import sys

CLASS_OF = {"k_ntt_b_ks_all": "ntt_ks_fused", "k_ks_baby_fused": "ks_baby_fused", "k_pmac_tma": "pmac",
            "k_intt_modup_fwd_a": "modup"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = float(r[col[name]].replace(",", ""))
        return v * UNIT.get(units[col[name]], 1.0)
    table, modup = {}, []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        cls = next((c for k, c in CLASS_OF.items() if k in name), None)
        if cls is None:
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        ms = float(r[col["gpu__time_duration.sum"]])
        if units[col["gpu__time_duration.sum"]] in ("us", "usecond"):
            ms /= 1e3
        e = {"kernel": name.split("(")[0].replace("void <unnamed>::", ""), "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd,
             "dram_write_bytes": wr, "duration_ms_under_ncu": ms,
             "dram_throughput_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
             "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
             "pipe_fmaheavy_cycles_active_pct": float(r[col["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
             "pipe_alu_cycles_active_pct": float(r[col["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]]),
             "warps_active_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
             "warp_instructions": float(r[col["smsp__inst_executed.sum"]].replace(",", "")),
             "registers_per_thread": int(float(r[col["launch__registers_per_thread"]])),
             "source": f"ncu --set full --clock-control none, one launch, C3 mat-vec (tools/profile_step.py), {rep}"}
        if cls == "modup":
            modup.append(e)
        else:
            table[cls] = e
    if modup:
        modup.sort(key=lambda e: e["duration_ms_under_ncu"])
        table["modup_small_launch"], table["modup_big_launch"] = modup[0], modup[-1]
        avg = dict(modup[-1])
        for k in ("dram_bytes_per_launch", "dram_read_bytes", "dram_write_bytes", "duration_ms_under_ncu"):
            avg[k] = sum(e[k] for e in modup) / len(modup)
        avg["note"] = "average of the two launches of a mat-vec (one decomposition for the baby steps, all giant groups at once)"
        table["modup"] = avg
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
