"""Developer script: per-source-line instruction counts / stall samples of one kernel from an ncu report.
usage: python tests/dev_ncu_lines.py report.ncu-rep <kernel regex> <mangled-name substring> [file.cu]
Joins `ncu --page source --csv` (SASS) with `nvdisasm -g` line info of the in-tree .so by instruction offset."""
import csv, io, os, re, subprocess, sys, collections, tempfile, glob
rep, kre, mangled = sys.argv[1:4]
src_filter = sys.argv[4] if len(sys.argv) > 4 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = rows[0]
ia, ii, isamp, isrc = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
sass = []
for r in rows[1:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        if sass: break   # only the first kernel instance
        continue
    sass.append((int(r[ia], 16), int(r[ii] or 0), int(r[isamp] or 0), r[isrc]))
base = sass[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "scri_b200", "libscrib200.so")], cwd=tmp, capture_output=True)
line_of = {}
for f in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", f], capture_output=True, text=True).stdout
    if mangled not in dis: continue
    sect = dis.split(f".text.{[w for w in re.findall(r'\.text\.(\S+):', dis) if mangled in w][0]}:")[1]
    cur = None
    for l in sect.splitlines():
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
        if m: line_of[int(m.group(1), 16)] = cur
        if l.startswith("//---") and line_of: break
    break
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for addr, n, s, _ in sass:
    k = line_of.get(addr - base)
    agg[k][0] += n; agg[k][1] += s; tot_i += n; tot_s += s
srcs = {}
print(f"total warp instructions {tot_i}, samples {tot_s}")
for k, (n, s) in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if k is None: print(f"  ?  inst {n} samples {s}"); continue
    if src_filter and k[0] != src_filter: txt = ""
    else:
        if k[0] not in srcs:
            p = glob.glob(os.path.join(root, "scri_b200", "csrc", k[0]))
            srcs[k[0]] = open(p[0]).read().splitlines() if p else []
        txt = srcs[k[0]][k[1] - 1].strip()[:90] if srcs[k[0]] and k[1] <= len(srcs[k[0]]) else ""
    print(f"{k[0]}:{k[1]:4d} inst {100*n/tot_i:5.1f}% samples {100*s/max(tot_s,1):5.1f}%  {txt}")
