#!/bin/bash
# A/B of bench.py under different environment settings: tools/ab_bench.sh VAR val1 val2 ... (prints value, msv ms, roofline)
VAR=$1; shift
for V in "$@"; do
  env $VAR=$V python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null > /tmp/ab_$V.json
  python - "$VAR" "$V" /tmp/ab_$V.json <<'PY'
import sys, json
d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
print(sys.argv[1], sys.argv[2], "reads/s %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "ms_msv %.1f" % d["stage_ms"]["ms_msv"],
      "frac %.3f" % d["roofline"]["frac"], "useful cells/clk/SM %.1f" % d["roofline"]["useful_cells_per_clk_per_sm"], d["stage_ms"])
PY
done
