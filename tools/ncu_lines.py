"""Join an ncu SASS source page with nvdisasm line info and aggregate warp-stall samples per
source line.  usage: python tools/ncu_lines.py <report.ncu-rep> <kernel mangled substring> <src file> [top]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, ksub, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "banzai_b200", "libbanzai_b200.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines_by_off = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if ksub not in dis:
        continue
    inside = False
    cur_any, cur_src = None, None
    for ln in dis.splitlines():
        if ln.startswith("//-----") and ".text." in ln:
            inside = ksub in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_any = (os.path.basename(m.group(1)), int(m.group(2)))
            if os.path.basename(m.group(1)) == os.path.basename(srcfile):
                cur_src = int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines_by_off[int(m.group(1), 16)] = (cur_src, cur_any, m.group(2).strip())

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# find the header row of the wanted kernel
start = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name" and ksub.split("ILi")[0][-12:] in "".join(r):
        start = i
        break
if start is None:
    start = 0
hdr = rows[start + 1]
idx = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[start + 2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
base = int(data[0][idx["Address"]], 16)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
per_line = {}
total = 0
for r in data:
    off = int(r[idx["Address"]], 16) - base
    n = int(r[idx["# Samples"]] or 0)
    total += n
    src = lines_by_off.get(off, (None, None, ""))[0]
    d = per_line.setdefault(src, {"n": 0, "inst": 0})
    d["n"] += n
    d["inst"] += int(r[idx["Instructions Executed"]] or 0)
    for s in stalls:
        d[s] = d.get(s, 0) + int(r[idx[s]] or 0)
src_lines = open(srcfile).read().splitlines()
print(f"total samples {total}")
for src, d in sorted(per_line.items(), key=lambda kv: -kv[1]["n"])[:top]:
    dom = sorted(stalls, key=lambda s: -d.get(s, 0))[:2]
    txt = src_lines[src - 1].strip()[:80] if src and src <= len(src_lines) else "?"
    print(f"{d['n'] / total * 100:5.1f}%  L{src}: {txt:80s} inst={d['inst']:>10d} "
          f"{dom[0][6:]}={d.get(dom[0], 0) / max(1, d['n']) * 100:.0f}% {dom[1][6:]}={d.get(dom[1], 0) / max(1, d['n']) * 100:.0f}%")
