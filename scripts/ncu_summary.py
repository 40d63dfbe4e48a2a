"""ncu report -> profiles/<tag>_ncu_full_raw.csv + profiles/<tag>_ncu_full_summary.json (run here, no GPU needed).
usage: python scripts/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep <tag>
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
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


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
    for rec in out:
        dram = (rec.get("dram__bytes_read.sum", 0) + rec.get("dram__bytes_write.sum", 0))
        print(f"{rec['Kernel Name'][:44]:44s} {rec.get('gpu__time_duration.sum', 0):8.1f} us  regs {rec.get('launch__registers_per_thread', 0):4.0f}  "
              f"warps {rec.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):5.1f} %  IPC {rec.get('sm__inst_executed.avg.per_cycle_active', 0):4.2f}  "
              f"L1/L2 hit {rec.get('l1tex__t_sector_hit_rate.pct', 0):4.0f}/{rec.get('lts__t_sector_hit_rate.pct', 0):4.0f} %  DRAM {dram:6.1f} MB  "
              f"inst {rec.get('smsp__inst_executed.sum', 0) / 1e6:6.1f} M  LDL/STL {rec.get('smsp__inst_executed_op_local_ld.sum', 0) / 1e6:.2f}/{rec.get('smsp__inst_executed_op_local_st.sum', 0) / 1e6:.2f} M  "
              f"{', '.join(rec['top_stalls'])}")


if __name__ == "__main__":
    main()
