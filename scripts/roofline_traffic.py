"""profiles/roofline_traffic.json from an ncu report: DRAM bytes per launch of the dominant kernel.
  python scripts/roofline_traffic.py gpurun_out/prof_attn_bwd_kernel_r02.ncu-rep attn_bwd_kernel profiles/r02_ncu_attn_bwd_full.txt"""
import csv, io, json, os, subprocess, sys
rep, kernel, cited = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m[u]
vals = []
for r in body:
    if kernel in r[ix["Kernel Name"]]:
        vals.append(sum(to_bytes(r[ix[m]], units[ix[m]]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
assert vals, f"no launch of {kernel} in {rep}"
out = {"kernel": kernel, "dram_bytes_per_launch": int(sum(vals) / len(vals)), "launches_in_capture": len(vals),
       "source": f"{cited}: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (scripts/profile_r02.sh, "
                 f"scripts/roofline_traffic.py; 16 images per launch = one training-step call)"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out))
