"""Condense one `ncu --set full` capture of k_ipm_solve into profiles/<name>.json.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/r1_ipm_ncu_summary.json "<note>"
"""
import csv, io, json, subprocess, sys

rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
SC = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def f(k, scale=True):
    u, v = m[k]
    x = float(v.replace(",", ""))
    return x * SC.get(u.split('/')[0], 1.0) if scale else x


rd, wr, ms = f("dram__bytes_read.sum"), f("dram__bytes_write.sum"), f("gpu__time_duration.sum")
d = {"source": "ncu --set full --clock-control none --import-source on -k regex:k_ipm_solve -c 1; " + note,
     "kernel": m["Kernel Name"][1], "grid": int(f("launch__grid_size")), "block": int(f("launch__block_size")),
     "registers_per_thread": int(f("launch__registers_per_thread")),
     "smem_dynamic_bytes": f("launch__shared_mem_per_block_dynamic"), "smem_static_bytes": f("launch__shared_mem_per_block_static"),
     "gpu_time_ms": ms, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
     "dram_gbs_under_ncu": (rd + wr) / ms / 1e6,
     "l2_hit_rate_pct": f("lts__t_sector_hit_rate.pct"), "l1_hit_rate_pct": f("l1tex__t_sector_hit_rate.pct"),
     "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
     "sm_throughput_pct": f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
     "warp_instructions": f("smsp__inst_executed.sum")}
for k in ("smsp__issue_active.avg.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"):
    if k in m:
        d[k] = f(k)
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))
