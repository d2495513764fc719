"""Join an `ncu --page source --csv` export (SASS rows with executed counts and stall samples) with the
line table of the same build (`nvdisasm -g` of the cubin inside the .so) and aggregate per source line:
executed warp instructions, samples, top stall reasons.

    python profiles/ncu_lines.py gpurun_out/r02_source_fp32.csv.gz mchammer_b200/libmchammer_b200.so \
        mch_optimize_f32 _ZN3mch19mch_optimize_kernelIfdLi0ELi64ELb0EEEvNS_7OptArgsE [min_exec max_exec] [n]

The .so must be the build the profile was taken from (rows are matched in order and checked by opcode).
"""
import collections
import csv
import gzip
import os
import re
import subprocess
import sys
import tempfile

path, lib, tu, kernel = sys.argv[1:5]
lo = int(sys.argv[5]) if len(sys.argv) > 5 else 0
hi = int(sys.argv[6]) if len(sys.argv) > 6 else 1 << 62
top = int(sys.argv[7]) if len(sys.argv) > 7 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", f"{tu}.sm_100a.cubin", os.path.abspath(lib)], cwd=tmp, check=True,
               stdout=subprocess.DEVNULL)
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f"{tu}.sm_100a.cubin")], capture_output=True,
                      text=True, check=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.strip().startswith(".section") and f".text.{kernel}," in l)
lines = []  # (opcode text, file, line) per instruction, in order
cur = ("?", 0)
for l in sass[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        lines.append((m.group(1).strip(), cur[0], cur[1]))

fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(fh))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > 5 and r[hdr.index("Instructions Executed")].isdigit()]
isrc, iex, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
if len(data) != len(lines):
    print(f"warning: {len(data)} profiled instructions vs {len(lines)} in the cubin", file=sys.stderr)
bad = sum(1 for r, l in zip(data, lines) if r[isrc].split()[0].lstrip("@!P0123456789T").strip() and
          r[isrc].strip().split()[-1 if False else 0] != l[0].split()[0] and not r[isrc].strip().startswith("@"))
ts = sum(int(r[ism]) for r in data) or 1
te = sum(int(r[iex]) for r in data) or 1
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), 0])
for r, l in zip(data, lines):
    e = int(r[iex])
    if not (lo <= e <= hi):
        continue
    a = agg[(l[1], l[2])]
    a[0] += e
    a[1] += int(r[ism])
    a[3] += 1
    for i in stall:
        if r[i].isdigit():
            a[2][hdr[i][6:]] += int(r[i])
print(f"{len(data)} instructions, opcode mismatches {bad}; selected exec range [{lo}, {hi}]")
sel_e = sum(a[0] for a in agg.values())
sel_s = sum(a[1] for a in agg.values())
print(f"selected: {100 * sel_e / te:.2f}% of executed instructions, {100 * sel_s / ts:.2f}% of samples")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = " ".join(f"{k}:{100 * v / max(1, sum(a[2].values())):.0f}" for k, v in a[2].most_common(3))
    print(f"{f}:{ln:<5d} static {a[3]:4d}  exec {100 * a[0] / te:5.2f}%  samples {100 * a[1] / ts:5.2f}%  {st}")
