set -x
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR scripts/h2d_bw_nranks.py > gpurun_out/h2d_n$N.json 2> gpurun_out/h2d_n$N.err
$TR scripts/h2d_bw_nranks.py --bind 0 > gpurun_out/h2d_n${N}_nobind.json 2>> gpurun_out/h2d_n$N.err
DVFE_GRAPHS=1 $TR bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/r2_n${N}_g1.json 2> gpurun_out/r2_n${N}_g1.err
DVFE_GRAPHS=0 $TR bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/r2_n${N}_g0.json 2> gpurun_out/r2_n${N}_g0.err
cat gpurun_out/h2d_n$N.json gpurun_out/h2d_n${N}_nobind.json
python - <<PY
import json
for g in (1,0):
    try:
        d=json.load(open("gpurun_out/r2_n${N}_g%d.json"%g)); print("graphs",g,"value",round(d["value"]),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]),"e2e ms",round(d["e2e"]["ms_per_step"],3), d["config"]["host_placement"])
    except Exception as e: print("fail",g,e)
PY
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|model name" | head -8
