#!/bin/bash
# Regenerates the tracked round-2 profile artefacts from the two captures a GPU run leaves in gpurun_out/:
#   gpurun_out/launches_r2_final.csv   ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 3000 --csv --log-file ... python bench.py --steps 4 --warmup 3 --no-cpu-baseline
#   gpurun_out/r2s2_full.ncu-rep       ncu --set full --clock-control none --import-source on -k regex:"k_lk_track|k_gftt_response|k_pyr_down|k_gftt_select" --launch-skip 21 -c 24 -o ... python bench.py --steps 1 --warmup 3 --groups 1 --no-cpu-baseline
# -> profiles/launches_r2.csv, traffic.json (stamped), ncu_r2_hotspots.txt, sass_r2_excerpts.txt and the tables for ncu_r2_summary.md (/tmp/r2_tables.md)
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/launches_r2_final.csv profiles/launches_r2.csv
python scripts/ncu_traffic.py gpurun_out/r2s2_full.ncu-rep > /dev/null
python scripts/ncu_round_summary.py profiles/launches_r2.csv gpurun_out/r2s2_full.ncu-rep > /tmp/r2_tables.md
ncu -i gpurun_out/r2s2_full.ncu-rep --page source --csv > /tmp/sass_r2s2.csv 2>/dev/null
python scripts/ncu_hotspots.py /tmp/sass_r2s2.csv 16 > /tmp/hot_all.txt
python scripts/sass_excerpts.py > profiles/sass_r2_excerpts.txt
python - <<'PY'
import json, re
tj = json.load(open("profiles/traffic.json"))
blocks = [("== " + b) if not b.startswith("==") else b for b in open("/tmp/hot_all.txt").read().split("\n== ")]
def wi(stage): return str(int(tj["_sm_l1"][stage]["warp_inst"]))
out = []
for w in (wi("lk_temporal"), wi("lk_stereo"), wi("gftt_response")):
    out += [b.rstrip() for b in blocks if ("warp instructions " + w) in b.split("\n")[0]][:1]
out += [b.rstrip() for b in blocks if re.search(r"k_gftt_select.*warp instructions (\d+)", b.split("\n")[0]) and
        int(re.search(r"warp instructions (\d+)", b.split("\n")[0]).group(1)) < 20000000][:1]
seen = set()
for b in blocks:
    m = re.search(r"k_pyr_down.*warp instructions (\d+)", b.split("\n")[0])
    if m and m.group(1) not in seen:
        seen.add(m.group(1)); out.append(b.rstrip())
open("profiles/ncu_r2_hotspots.txt", "w").write("\n".join(out) + "\n")
for s in ("lk_temporal", "lk_stereo", "gftt_response"):
    b = [b for b in out if ("warp instructions " + wi(s)) in b.split("\n")[0]]
    print(s, tj["_sm_l1"][s], "\n   ", b[0].split("\n")[1].strip() if b else "")
print("traffic", {k: v for k, v in tj.items() if not k.startswith("_")}, tj["_kernel_sources_sha"])
PY
