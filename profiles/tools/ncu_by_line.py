#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line.
usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
Joins `ncu --page source --print-source sass --csv` (per-instruction counters) with `nvdisasm -g` line info of the same
cubin (the .so that was profiled) and prints, per source line: warp instructions, thread instructions, avg active lanes,
stall samples."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = ""   # the library holds one cubin per translation unit: take the one that defines the kernel
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        d = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, f)]).decode()
        if re.search(r"\.text\.\S*" + re.escape(kern), d):
            dis = d
            break
addr2line, cur, infn = {}, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = kern in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m: addr2line[int(m.group(1), 16)] = cur
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], stderr=subprocess.DEVNULL).decode()
blocks = re.split(r'(?m)^"Kernel Name",', out)
want = kern.replace("ILb1", "<(bool)1>").replace("ILb0", "<(bool)0>")
agg = collections.defaultdict(lambda: [0, 0, 0])
done = False
for b in blocks[1:]:
    name, rest = b.split("\n", 1)
    if want not in name and kern not in name: continue
    if done: break
    done = True
    rd = csv.DictReader(io.StringIO(rest))
    base = None
    for r in rd:
        try: a = int(r["Address"], 16)
        except Exception: continue
        if base is None: base = a
        key = addr2line.get(a - base) or ("?", 0)
        ie, te = int(r["Instructions Executed"] or 0), int(r["Thread Instructions Executed"] or 0)
        ss = int(r["# Samples"] or 0)
        agg[key][0] += ie; agg[key][1] += te; agg[key][2] += ss
tot = [sum(v[i] for v in agg.values()) for i in range(3)]
print(f"total warp-inst {tot[0]:,} thread-inst {tot[1]:,} avg lanes {tot[1]/max(tot[0],1):.2f} samples {tot[2]:,}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), v in agg.items():
    for i in range(3): byfile[f][i] += v[i]
for f, v in sorted(byfile.items(), key=lambda x: -x[1][2]):
    print(f"  {f:28s} warp-inst {100*v[0]/tot[0]:5.1f}%  samples {100*v[2]/max(tot[2],1):5.1f}%  lanes {v[1]/max(v[0],1):5.2f}")
print("top lines by stall samples:")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    print(f"  {f}:{l:<5d} warp-inst {100*v[0]/tot[0]:5.2f}%  samples {100*v[2]/max(tot[2],1):5.2f}%  lanes {v[1]/max(v[0],1):5.2f}")
