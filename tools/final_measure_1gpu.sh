set -x
mkdir -p gpurun_out/f
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/f/pytest_gpu.txt 2>&1; tail -3 gpurun_out/f/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/f/r2_bench_1gpu_r18_50k_v6.json 2> gpurun_out/f/bench_50k.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f/r2_bench_reference_arm_v6.json 2>/dev/null
python bench.py --workload r18_2k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f/r2_bench_1gpu_r18_2k_v6.json 2>/dev/null
python bench.py --workload r152_highreg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f/r2_bench_1gpu_r152_highreg_v6.json 2>/dev/null
python bench.py --workload r18_sgd --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f/r2_bench_1gpu_r18_sgd_v6.json 2>/dev/null
python bench.py --workload r18_500k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f/r2_bench_1gpu_r18_500k_v6.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f/smoke.txt 2>&1; tail -2 gpurun_out/f/smoke.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f/r2_launches_g8_mb128_v4.csv python tools/profile_step.py 18 128 split 8 > gpurun_out/f/profile_step.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_kernel -c 30 -o gpurun_out/f/r2_conv_full_v4 python tools/profile_step.py 18 128 split 8 > gpurun_out/f/ncu_full.log 2>&1
ncu -i gpurun_out/f/r2_conv_full_v4.ncu-rep --page raw --csv > gpurun_out/f/r2_conv_full_v4_raw.csv 2>/dev/null
rm -f gpurun_out/f/r2_conv_full_v4.ncu-rep
for f in gpurun_out/f/r2_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d.get("value",0)), round(d.get("e2e",{}).get("value",0)), d.get("clocks"), d.get("check",{}).get("ok"), d.get("gpu_launches"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
