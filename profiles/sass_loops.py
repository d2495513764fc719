"""List the loops (backward branches) of one kernel in a built library with their instruction mix.

    python profiles/sass_loops.py mchammer_b200/libmchammer_b200.so mch_optimize_kernelIfdLi0ELi128
"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
body = next(b for b in blocks if pat in b.split("\n", 1)[0])
ins = []
for line in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
print(len(ins), "instructions")
code = {"FADD": "a", "FMUL": "m", "FFMA": "f", "FFMA2": "F", "FADD2": "A", "FMUL2": "M", "MUFU": "R", "LDS": "L", "STS": "S", "BRA": "B",
        "LDL": "!", "STL": "!", "DADD": "d", "DFMA": "d", "DMUL": "d", "BAR": "|"}
for i, (a, text) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", text)
    if not m:
        continue
    target = int(m.group(1), 16)
    if target in addr_index and target < a:
        j = addr_index[target]
        seg = [t for _, t in ins[j:i + 1]]
        ops = [(t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for t in seg]
        n_mufu = ops.count("MUFU")
        if n_mufu >= 2:
            s = "".join(code.get(o, "u") for o in ops)
            print(f"loop @{target:#x}..{a:#x}: {len(seg)} instr, {n_mufu} MUFU, {len(seg) / n_mufu:.2f} instr/MUFU")
            print("   ", s)
