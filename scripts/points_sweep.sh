#!/bin/bash
# 1280x720 stereo at 150 / 400 / 1000 points per frame (north_star: ">= 50x ... at 150-1000 points"): our arm and the
# reference arm (all host cores), 64 streams
for cfg in "150 30" "400 25" "1000 10"; do
  set -- $cfg
  python bench.py --max-cnt $1 --min-dist $2 --steps 100 --no-cpu-baseline --e2e-groups $([ $1 -ge 1000 ] && echo 4 || echo 1) > /tmp/a.json 2>/dev/null
  python bench.py --impl reference --max-cnt $1 --min-dist $2 --steps 8 --warmup 3 > /tmp/r.json 2>/dev/null
  python - $1 $2 <<PY
import json, sys
a, r = json.load(open("/tmp/a.json")), json.load(open("/tmp/r.json"))
print("max_cnt %s min_dist %s: value %.0f e2e %.0f reference %.0f (%d cores)  e2e/ref %.1fx  pts/step %d" % (
    sys.argv[1], sys.argv[2], a["value"], a["e2e"]["value"], r["value"], r["cpu_baseline"]["cores"],
    a["e2e"]["value"] / r["value"], a["config"]["tracked_points_per_step"]))
PY
done
