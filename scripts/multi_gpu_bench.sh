#!/bin/bash
# usage: gpurun --gpus N -- 'bash scripts/multi_gpu_bench.sh N'   -- the driver's N-GPU launch of bench.py, graphs on and off, + H2D probe
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out/evidence; mkdir -p $O
for g in ${GRAPHS:-1 0}; do
  DVFE_GRAPHS=$g $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c5_n${N}_g$g.json 2> $O/bench_c5_n${N}_g$g.err
done
$TR scripts/h2d_bw_nranks.py > $O/h2d_n$N.json 2> $O/h2d_n$N.err
python - <<PY
import json
for g in [int(x) for x in "${GRAPHS:-1 0}".split()]:
    try:
        d = json.loads(open("$O/bench_c5_n${N}_g%d.json" % g).read().strip().splitlines()[-1])
        print("N=$N graphs", g, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              "e2e roofline", d["e2e"].get("roofline", {}).get("frac"))
    except Exception as e:
        print("fail", g, e)
print(open("$O/h2d_n$N.json").read().strip()[-400:])
PY
