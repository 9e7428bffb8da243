#!/bin/bash
# bench every library under variants/ (built with different -D switches): value, e2e and the stage split
for f in dynamic_vins_b200/libdvfe.so variants/*.so; do
  DVFE_LIB=$PWD/$f python bench.py --no-cpu-baseline --steps ${STEPS:-60} > /tmp/v.json 2>/dev/null
  python - "$f" <<PY
import json, sys
d = json.load(open("/tmp/v.json"))
s = d["stage_ms"]
print(sys.argv[1], "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), " ".join("%s %.3f" % (k, v) for k, v in s.items() if v > 0.05))
PY
done
