"""Break an `ncu --page source --csv` export down by code region (runs of instructions with the same
execution count): share of warp-time samples, share of executed instructions, top stall reasons, mix.

    python profiles/ncu_regions.py gpurun_out/r02_source_fp32.csv.gz [n_regions]
"""
import collections
import csv
import gzip
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(fh))
hdr = next(r for r in rows if r and r[0] == "Address")
data = rows[rows.index(hdr) + 1:]
isrc, iex, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in data if len(r) > iex and r[iex].isdigit()]
ts = sum(int(r[ism]) for r in data) or 1
te = sum(int(r[iex]) for r in data) or 1
regions, cur = [], None
for idx, r in enumerate(data):
    e = int(r[iex])
    if cur and cur["e"] == e:
        cur["rows"].append(r)
    else:
        cur = {"e": e, "start": idx, "rows": [r]}
        regions.append(cur)
for g in regions:
    g["s"] = sum(int(r[ism]) for r in g["rows"])
regions.sort(key=lambda g: -g["s"])
print(f"{len(data)} instructions, {ts} samples")
for g in regions[:top]:
    agg, ops = collections.Counter(), collections.Counter()
    for r in g["rows"]:
        for i in stall:
            if r[i].isdigit():
                agg[hdr[i][6:]] += int(r[i])
        f = r[isrc].split()
        ops[(f[1] if f[0].startswith("@") else f[0]).split(".")[0]] += 1
    tot = sum(agg.values()) or 1
    n = len(g["rows"])
    print(f"@{g['start']:5d} n={n:4d} exec={g['e']:>11d} samples={100 * g['s'] / ts:5.2f}% instr={100 * g['e'] * n / te:5.2f}% | "
          + " ".join(f"{k}:{100 * v / tot:.0f}" for k, v in agg.most_common(4)) + " | "
          + " ".join(f"{k}{v}" for k, v in ops.most_common(6)))
