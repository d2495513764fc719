"""Summarise an .ncu-rep (raw + source pages) into the few numbers DESIGN.md quotes.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--json profiles/rNN_ncu_full.json]

The JSON (per-launch DRAM bytes, duration, occupancy, pipe utilisation) is what bench.py reads
for `roofline.traffic`.
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    summary = {}
    for row in raw[2:]:
        print("kernel:", row[hdr.index("Kernel Name")][:90])
        summary = {"kernel": row[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                print(f"  {k:68s} {row[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
                try:
                    summary[k] = {"value": float(row[hdr.index(k)].replace(",", "")), "unit": units[hdr.index(k)]}
                except ValueError:
                    pass

    def to_bytes(key):
        item = summary.get(key)
        if not item:
            return None
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(item["unit"], 1)
        return item["value"] * scale

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    if rd is not None and wr is not None:
        summary["dram_bytes_per_launch"] = rd + wr
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
            json.dump(summary, fh, indent=1)
    src = page(rep, "source")
    hdr = next(r for r in src if r and r[0] == "Address")
    rows = src[src.index(hdr) + 1:]
    isrc, iex, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows if len(r) > iex and r[iex].isdigit()]
    tot = sum(int(r[iex]) for r in data)
    ts = sum(int(r[ism]) for r in data)
    byop, reasons = collections.Counter(), collections.Counter()
    for r in data:
        f = r[isrc].split()
        op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
        byop[op] += int(r[iex])
        for i in stall:
            if r[i].isdigit():
                reasons[hdr[i]] += int(r[i])
    print("instruction mix (% of warp instructions):",
          ", ".join(f"{op} {100 * c / tot:.1f}" for op, c in byop.most_common(12)))
    s = sum(reasons.values())
    print("stall samples (%):", ", ".join(f"{k[6:]} {100 * v / s:.1f}" for k, v in reasons.most_common(8)))
    mufu = byop.get("MUFU", 0)
    if mufu:
        print(f"warp instructions per MUFU (= per 32 pair terms): {tot / mufu:.2f}")
    top = sorted(data, key=lambda r: -int(r[ism]))[:8]
    for r in top:
        print(f"  {100 * int(r[ism]) / ts:5.2f}% samples  exec {r[iex]:>12s}  {r[isrc].strip()[:60]}")


if __name__ == "__main__":
    main()
