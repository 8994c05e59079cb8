"""Dev tool: per-kernel DRAM traffic and time from an ncu report -> JSON (what bench.py's `roofline.traffic` reads).
usage: python dev/dev_ncu_traffic.py report.ncu-rep out.json"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
ki = h.index("Kernel Name")


def val(r, name):
    i = h.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1)
    return v * scale


res = {}
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("scrib200::", "").replace("(int)", "")
    res[name] = {
        "dram_bytes_per_launch": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
        "gpu_time_us_under_ncu": val(r, "gpu__time_duration.sum"),
    }
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
