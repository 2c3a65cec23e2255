"""Summarise one `ncu --set full` capture of knn2_kernel into profiles/knn2_traffic.json (read by bench.py for
roofline.traffic) and a details CSV.  Usage: python scripts/ncu_traffic.py gpurun_out/<tag>_knn2.ncu-rep <tag> <pairs_in_launch>"""
import csv, io, json, subprocess, sys
rep, tag, pairs = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(key):
    u, v = d[key]
    x = float(v.replace(",", ""))
    return x * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


out = {
    "kernel": "knn2_kernel",
    "source": f"profiles/{tag}_knn2_ncu_details.csv (ncu --set full, one launch = {pairs} pairs of config 3)",
    "dram_bytes_read": num("dram__bytes_read.sum"),
    "dram_bytes_write": num("dram__bytes_write.sum"),
    "pairs_in_launch": pairs,
    "algorithmic_bytes_per_launch": pairs * (2 * 10000 * 128 + 10000 * 16),
    "gpu_time_ms": num("gpu__time_duration.sum"),
    "tensor_pipe_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": int(num("launch__registers_per_thread")),
}
out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
json.dump(out, open("profiles/knn2_traffic.json", "w"), indent=1)
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
open(f"profiles/{tag}_knn2_ncu_details.csv", "w").write(det)
print(json.dumps(out, indent=1))
