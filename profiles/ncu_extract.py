#!/usr/bin/env python3
"""Summarise an `ncu --set full` capture (read here, without a GPU):

    ncu -i gpurun_out/r02/prof_hits.ncu-rep --page raw --csv > profiles/r02/ncu_full_hits_raw.csv
    python profiles/ncu_extract.py profiles/r02/ncu_full_hits_raw.csv [--traffic-json motifs genome_bp rate]

Prints one block per profiled launch (duration, DRAM bytes, shared-memory wavefronts, pipe utilisation,
issue rate, stall reasons per issue) and, with --traffic-json, writes profiles/r02/ncu_hits_traffic.json
(the `roofline.traffic` bench.py reports) from the LAST scan_hits_kernel launch in the file."""
import csv
import json
import os
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "smem_ld_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smem_ld_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pipe_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex_data_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_pct"),
    ("sm__warps_active.avg.per_cycle_active", "warps_per_sm"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_per_smsp"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "local_ld_requests"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col.get("Kernel Name")
    out = []
    for r in rows[2 if units else 1:]:
        if len(r) < len(hdr):
            continue
        rec = {"kernel": r[name_i].split("(")[0] if name_i is not None else "?"}
        for k, short in KEYS:
            if k in col and r[col[k]] not in ("", "n/a"):
                try:
                    rec[short] = float(r[col[k]].replace(",", ""))
                    if units:
                        rec.setdefault("_units", {})[short] = units[col[k]]
                except ValueError:
                    pass
        stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(r[i]) for h, i in col.items()
                  if "warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")}
        rec["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        out.append(rec)
    for rec in out:
        print(json.dumps(rec))
    if "--traffic-json" in sys.argv:
        i = sys.argv.index("--traffic-json")
        motifs, genome_bp, rate = int(sys.argv[i + 1]), int(sys.argv[i + 2]), float(sys.argv[i + 3])
        scans = [r for r in out if "scan_hits_kernel" in r["kernel"] and r.get("duration", 0) > 0]
        last = max(scans, key=lambda r: r.get("duration", 0))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        u = last.get("_units", {})
        t = {"motifs": motifs, "genome_bp": genome_bp, "rate": rate, "kernel": "scan_hits_kernel",
             "dram_bytes_read": last["dram_read"] * scale.get(u.get("dram_read", "byte"), 1.0),
             "dram_bytes_write": last["dram_write"] * scale.get(u.get("dram_write", "byte"), 1.0),
             "source": os.path.basename(path), "how": "ncu --set full --clock-control none, one launch"}
        dst = os.path.join(os.path.dirname(os.path.abspath(path)), "ncu_hits_traffic.json")
        json.dump(t, open(dst, "w"), indent=1)
        print("wrote", dst, t)


if __name__ == "__main__":
    main()
