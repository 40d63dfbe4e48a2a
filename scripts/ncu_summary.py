"""ncu report -> profiles/<tag>_ncu_full_raw.csv + profiles/<tag>_ncu_full_summary.json (run here, no GPU needed).
usage: python scripts/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep <tag> [--kernels-json]
--kernels-json also (re)writes profiles/ncu_kernels.json, the per-kernel figures bench.py reports as `roofline.measured`
(merged over calls: a later report only replaces the kernels it contains).
The summary keeps, per captured launch, the metrics DESIGN.md and bench.py's `roofline.measured` quote."""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors.sum", "l1tex__t_bytes.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def short_name(full):
    """'void <unnamed>::k_shadow<8, 1>(DScene, DFrame)' -> 'k_shadow'"""
    n = full.split("(")[0].split("::")[-1]
    return n.split("<")[0].strip()


def kernels_json(out, tag):
    import os
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    path = "profiles/ncu_kernels.json"
    doc = json.load(open(path)) if os.path.exists(path) else {"kernels": {}}
    by = {}
    for rec in out:
        by.setdefault(short_name(rec["Kernel Name"]), []).append(rec)
    for name, recs in by.items():
        recs.sort(key=lambda r: r.get("gpu__time_duration.sum", 0.0))
        r = recs[len(recs) // 2]                       # the launch with the median duration
        us = r.get("gpu__time_duration.sum", 0.0)
        unit = r.get("gpu__time_duration.sum [unit]", "us")
        us = us / 1000.0 if unit in ("ns", "nsecond") else us * 1000.0 if unit in ("ms", "msecond") else us

        def mb(key):
            v, u = r.get(key, 0.0), r.get(key + " [unit]", "byte")
            return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        dram = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
        l2 = mb("lts__t_bytes.sum") or 32.0 * r.get("lts__t_sectors.sum", 0.0)          # 32-byte sectors
        dram_gbs = dram / (us * 1e-6) / 1e9 if us > 0 else 0.0
        issue = r.get("smsp__issue_active.avg.pct_of_peak_sustained_active")
        doc["kernels"][name] = {
            "kernel": r["Kernel Name"], "kernel_us": us, "launches_captured": len(recs), "registers": r.get("launch__registers_per_thread"),
            "dram_bytes": dram, "dram_gbs": dram_gbs, "dram_frac": dram_gbs / hbm, "l2_gbs": l2 / (us * 1e-6) / 1e9 if us > 0 and l2 else None,
            "issue_slot_busy_pct": issue, "ipc": r.get("sm__inst_executed.avg.per_cycle_active"),
            "warps_active_pct": r.get("sm__warps_active.avg.pct_of_peak_sustained_active"), "warp_instructions": r.get("smsp__inst_executed.sum"),
            "l1_hit_pct": r.get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": r.get("lts__t_sector_hit_rate.pct"),
            "tensor_pipe_pct": r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "local_memory_sectors": (r.get("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", 0.0) or 0.0) + (r.get("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", 0.0) or 0.0),
            "top_stalls": r.get("top_stalls"),
            "bound": "issue+latency (dependent node fetches; DRAM %.0f %% of the measured copy bandwidth)" % (100.0 * dram_gbs / hbm) if dram_gbs / hbm < 0.5 else "hbm",
            "capture": tag,
        }
    doc["_source"] = ("ncu --set full --clock-control none --import-source on, each kernel alone, serialised, cold cache; the launch with the "
                      "median duration per kernel; summaries profiles/<capture>_ncu_full_summary.json; dram_frac against MEASURED_PEAKS.json hbm_gbs = %.1f" % hbm)
    json.dump(doc, open(path, "w"), indent=1)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(f"profiles/{tag}_ncu_full_raw.csv", "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        rec = {"Kernel Name": d["Kernel Name"], "ID": d.get("ID")}
        for k, u in zip(head, units):
            if k in KEEP or (k.startswith(STALL) and k.endswith("_per_issue_active.ratio")):
                try:
                    rec[k] = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if u:
                    rec[k + " [unit]"] = u
        stalls = sorted(((v, k[len(STALL):-len("_per_issue_active.ratio")]) for k, v in rec.items() if k.startswith(STALL) and isinstance(v, float)), reverse=True)
        rec["top_stalls"] = [f"{name} {v:.2f}" for v, name in stalls[:4]]
        out.append(rec)
    json.dump(out, open(f"profiles/{tag}_ncu_full_summary.json", "w"), indent=1)
    if "--kernels-json" in sys.argv:
        kernels_json(out, tag)
    for rec in out:
        dram = (rec.get("dram__bytes_read.sum", 0) + rec.get("dram__bytes_write.sum", 0))
        print(f"{rec['Kernel Name'][:44]:44s} {rec.get('gpu__time_duration.sum', 0):8.1f} us  regs {rec.get('launch__registers_per_thread', 0):4.0f}  "
              f"warps {rec.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):5.1f} %  IPC {rec.get('sm__inst_executed.avg.per_cycle_active', 0):4.2f}  "
              f"L1/L2 hit {rec.get('l1tex__t_sector_hit_rate.pct', 0):4.0f}/{rec.get('lts__t_sector_hit_rate.pct', 0):4.0f} %  DRAM {dram:6.1f} MB  "
              f"inst {rec.get('smsp__inst_executed.sum', 0) / 1e6:6.1f} M  LDL/STL {rec.get('smsp__inst_executed_op_local_ld.sum', 0) / 1e6:.2f}/{rec.get('smsp__inst_executed_op_local_st.sum', 0) / 1e6:.2f} M  "
              f"{', '.join(rec['top_stalls'])}")


if __name__ == "__main__":
    main()
