"""Print the few headline numbers of a bench.py JSON line read from stdin."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d["roofline"]
print(sys.argv[1] if len(sys.argv) > 1 else "", "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"],
      "pairs/s %.3e" % d["pair_evals_per_s"], "frac %.3f" % r["frac"], "peak %.1f" % r["peak"],
      "ms %.2f" % d["ms_per_step"], "sm_mhz", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
