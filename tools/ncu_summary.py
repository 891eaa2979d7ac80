"""Development aid: text + json summary of the counters the roofline argument uses, from ncu --set full reports.
usage: ncu_summary.py out_prefix report.ncu-rep[:kernel-regex[:batch]] ...   -> <out_prefix>_summary.txt, <out_prefix>_metrics.json"""
import csv, io, json, subprocess, sys

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sass__inst_executed_register_spilling", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}

prefix = sys.argv[1]
txt, js = [], {}
for spec in sys.argv[2:]:
    parts = spec.split(":")
    rep, kern, batch = parts[0], (parts[1] if len(parts) > 1 else ""), (int(parts[2]) if len(parts) > 2 else None)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r)); u = dict(zip(head, units))
        if kern and kern not in d.get("Kernel Name", ""): continue
        txt.append("Kernel Name = " + d["Kernel Name"])
        for m in METRICS:
            if m in d and d[m] != "": txt.append(f"{m} = {d[m]} {u[m]}")
        stalls = sorted(((float(d[k].replace(",", "")), k) for k in head if k.startswith("smsp__average_warp_latency_issue_stalled") or k.startswith("smsp__average_warps_issue_stalled")
                         if d[k] not in ("", "n/a")), reverse=True)[:8] if False else []
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("b200::", "").replace("(int)", "").replace(", 0>", ">")
        f = lambda m: float(d[m].replace(",", "")) * UNIT.get(u[m], 1.0)
        js[name] = {"batch": batch, "dram_bytes_read": f("dram__bytes_read.sum"), "dram_bytes_write": f("dram__bytes_write.sum"),
                    "gpu_time_ms": f("gpu__time_duration.sum"),
                    "fp64_pipe_active_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
                    "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"])}
        txt.append("----")
open(prefix + "_summary.txt", "w").write("\n".join(txt) + "\n")
json.dump(js, open(prefix + "_metrics.json", "w"), indent=1)
print("\n".join(txt))
