# usage: bash tools/final_measure_ngpu.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out/f
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu > gpurun_out/f/pytest_multigpu.txt 2>&1; tail -3 gpurun_out/f/pytest_multigpu.txt; fi
for wl in r18_50k r18_500k; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f/r2_bench_${N}gpu_${wl}_v6.json 2> gpurun_out/f/bench_${N}gpu_${wl}.err
  python - gpurun_out/f/r2_bench_${N}gpu_${wl}_v6.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d.get("value",0)), round(d.get("e2e",{}).get("value",0)), d.get("ms_per_step"), d.get("clocks"), d.get("check",{}).get("ok"), d.get("gpu_launches"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
