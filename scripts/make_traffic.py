"""profiles/<tag>_traffic.json: DRAM bytes per launch of the four kernels of a config-2 step, from the committed
`ncu --set full` capture (gpurun_out/<tag>_kernels.ncu-rep).  bench.py reads it for `roofline.traffic` / `frac_dram`."""
import csv
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 8513610
import os
rep = f"gpurun_out/{tag}_step_kernels.ncu-rep" if os.path.exists(f"gpurun_out/{tag}_step_kernels.ncu-rep") else f"gpurun_out/{tag}_kernels.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
names = {"march": "march_kernel", "expand": "expand_runs_vec_kernel", "composite_fwd": "composite_fwd_hot_kernel",
         "composite_bwd": "composite_bwd_hot_kernel"}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"n_samples": n_samples, "source": f"ncu --set full --clock-control none, {rep} (scripts/profile_kernels.py)"}
for key, kn in names.items():
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kn in d["Kernel Name"]:
            u = dict(zip(hdr, units))
            out[key] = {"dram_read": float(d["dram__bytes_read.sum"].replace(",", "")) * scale[u["dram__bytes_read.sum"]],
                        "dram_write": float(d["dram__bytes_write.sum"].replace(",", "")) * scale[u["dram__bytes_write.sum"]],
                        "gpu_time_us": float(d["gpu__time_duration.sum"].replace(",", ""))}
            break
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(json.dumps(out))
