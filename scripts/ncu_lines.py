"""Attribute the warp-stall samples of an ncu report to source lines (needs -lineinfo).
usage: ncu -i X.ncu-rep --page source --csv > /tmp/src.csv
       python scripts/ncu_lines.py /tmp/src.csv mdbn_b200/csrc/build/skinny_i10.o ILi10ELb1 40 [skinny_kernel.cuh]
(report csv, object with the kernel, kernel instantiation tag, rows to print, source file to attribute to)"""
import re, csv, collections, sys, subprocess, tempfile, os, glob
src_csv, obj, tag, nrows = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
fname = sys.argv[5] if len(sys.argv) > 5 else "skinny_kernel.cuh"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line = None; off2line = {}; infn = False
for l in dis.split("\n"):
    if l.lstrip().startswith(".section") and ".text." in l:
        infn = tag in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]*)", line (\d+)', l)
    if m:
        line = int(m.group(2)) if m.group(1).endswith(fname) else -int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = line
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]
base = int(data[0][0], 16)
si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
cols = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_not_selected", "stall_selected",
        "stall_math", "stall_mio", "stall_branch_resolving", "stall_lg", "stall_membar", "stall_dispatch"]
ci = [hdr.index(c) for c in cols]
agg = collections.defaultdict(lambda: [0, 0] + [0] * len(cols))
for r in data:
    if r and r[0] in ("Kernel Name", "Address"):
        if r[0] == "Kernel Name":
            break
        continue
    if len(r) < len(hdr):
        continue
    ln = off2line.get(int(r[0], 16) - base, -99999)
    a = agg[ln]; a[0] += int(r[si]); a[1] += int(r[ie])
    for k, i in enumerate(ci):
        a[2 + k] += int(r[i])
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print("total samples", tot, "warp-instructions", toti)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "mdbn_b200", "csrc", fname)).read().split("\n")
# per-region summary (25-line buckets) then the hottest lines
reg = collections.defaultdict(lambda: [0, 0])
for ln, a in agg.items():
    key = (ln // 25 * 25) if ln > 0 else -1
    reg[key][0] += a[0]; reg[key][1] += a[1]
print("--- regions (source line bucket: samples share, instruction share)")
for key, (s, i) in sorted(reg.items()):
    if s > 0.01 * tot:
        print(f"  {key:5d}: {100*s/tot:5.1f}%  inst {100*i/toti:5.1f}%")
print("--- hottest lines")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:nrows]:
    s = src[ln - 1].strip()[:70] if ln > 0 else "?"
    print(f"{ln:5d} {a[0]:6d} {100*a[0]/tot:5.1f}% inst={a[1]:9d} " + " ".join(f"{c[6:10]}={v}" for c, v in zip(cols, a[2:]) if v > 0.1 * a[0]) + " | " + s)
