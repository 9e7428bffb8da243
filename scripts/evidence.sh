#!/bin/bash
# Every single-GPU number quoted in DESIGN.md / README.md, in one run on a GPU box; JSON lines go to gpurun_out/evidence/ (copied to
# profiles/evidence_<round>/ by hand).  usage: gpurun --timeout 1500 -- 'bash scripts/evidence.sh'
O=gpurun_out/evidence; mkdir -p $O
python bench.py --steps 200 --warmup 10 > $O/bench_c5.json 2> $O/bench_c5.err
python bench.py --impl reference --steps 10 --warmup 3 > $O/ref_c5.json 2>> $O/bench_c5.err
for ms in 4 8; do python bench.py --steps 100 --motion-scale $ms --no-cpu-baseline > $O/bench_c5_churn$ms.json 2>/dev/null; done
for wl in c1_euroc_mono c2_kitti_stereo c4_hd_stereo; do
  g=1; [ $wl = c4_hd_stereo ] && g=4; [ $wl = c1_euroc_mono ] && g=2
  python bench.py --workload $wl --steps 100 --e2e-groups $g > $O/bench_$wl.json 2>/dev/null
  python bench.py --impl reference --workload $wl --steps 6 --warmup 2 > $O/ref_$wl.json 2>/dev/null
done
python bench.py --workload c3_zed_dynamic --steps 60 > $O/bench_c3_zed_dynamic.json 2>/dev/null
python bench.py --impl reference --workload c3_zed_dynamic --steps 4 --warmup 2 > $O/ref_c3_zed_dynamic.json 2>/dev/null
for cfg in "150 30" "1000 10"; do set -- $cfg
  python bench.py --max-cnt $1 --min-dist $2 --steps 100 --no-cpu-baseline --e2e-groups $([ $1 -ge 1000 ] && echo 4 || echo 1) > $O/bench_c5_pts$1.json 2>/dev/null
  python bench.py --impl reference --max-cnt $1 --min-dist $2 --steps 6 --warmup 2 > $O/ref_c5_pts$1.json 2>/dev/null
done
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob("gpurun_out/evidence/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-34s value %9.0f  e2e %9.0f  %s" % (os.path.basename(f), d["value"], d["e2e"]["value"],
              ("cores %d kind %s" % (d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])) if d.get("impl") == "reference" else
              ("ms/step %.3f parity %s" % (d["ms_per_step"], d.get("parity", {}).get("max_px_err_vs_ref_cpu")))))
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e)
PY
