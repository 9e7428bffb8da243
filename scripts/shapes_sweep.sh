#!/bin/bash
# the other BASELINE.json config shapes (C1, C2, C4): our arm and the reference arm on the same box, 64 streams
for wl in c1_euroc_mono c2_kitti_stereo c4_hd_stereo; do
  g=1; [ $wl = c4_hd_stereo ] && g=4; [ $wl = c1_euroc_mono ] && g=2   # e2e groups: 1 when PCIe-bound, more when upload ~ compute
  python bench.py --workload $wl --steps 100 --e2e-groups $g > /tmp/a.json 2>/dev/null
  python bench.py --impl reference --workload $wl --steps 6 --warmup 2 > /tmp/r.json 2>/dev/null
  python - $wl <<PY
import json, sys
a, r = json.load(open("/tmp/a.json")), json.load(open("/tmp/r.json"))
print("%s: value %.0f e2e %.0f reference %.0f (%d cores) e2e/ref %.1fx max_px_err %.2e ids_equal %s single %.3f ms" % (
    sys.argv[1], a["value"], a["e2e"]["value"], r["value"], r["cpu_baseline"]["cores"], a["e2e"]["value"] / r["value"],
    a["parity"]["max_px_err_vs_ref_cpu"], a["parity"]["ids_and_stereo_bits_equal"], a["single_stream"]["ms_per_frame_median"]))
PY
done
