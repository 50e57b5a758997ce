# dram bytes per launch of the hot kernels from an ncu summary (tools/ncu_summary.py output) -> profiles/<tag>_traffic.json
import json, re, sys
txt = open(sys.argv[1]).read()
out = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, 16384^2 fine level, one B200 (%s)" % sys.argv[1], "grid": [16384, 16384]}
names = {"k_mg_march_pre": "fused_pre", "k_mg_march_post": "fused_post", "k_pcg_matvec_march": "pcg_matvec", "k_update_xr": "update_xr", "k_muscl_x": "sweep_x", "k_muscl_y": "sweep_y", "k_build_links_march": "build_links", "k_velocity_march": "velocity"}
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
for blk in txt.split("== ")[1:]:
    head = blk.split("\n")[0]
    key = next((v for k, v in names.items() if k in head), None)
    if key is None or key in out: continue
    r = re.search(r"dram__bytes_read.sum\s+([\d.]+) (\w+)", blk); w = re.search(r"dram__bytes_write.sum\s+([\d.]+) (\w+)", blk)
    out[key] = float(r.group(1)) * unit[r.group(2)] + float(w.group(1)) * unit[w.group(2)]
import subprocess
try:
    out["commit"] = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], text=True).strip()
except Exception:
    pass
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
