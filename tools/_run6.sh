set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_env_gpu.py tests/test_mcts_gpu.py -q -x > gpurun_out/r2f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_tests.log
tail -3 gpurun_out/r2f_tests.log
for occ in 16 20 24; do
QZ_STUCK_OCC=$occ timeout 300 python bench.py --steps 8 --warmup 5 --no-az --no-kernels --no-parity --no-cpu-baseline --games-plies 0 > gpurun_out/r2f_bench_occ$occ.json 2> gpurun_out/r2f_bench_occ$occ.err
done
QZ_STUCK_OCC=24 timeout 300 python -m pytest tests/test_env_gpu.py -q -x -k "rollout or stuck" > gpurun_out/r2f_tests_occ24.log 2>&1; tail -2 gpurun_out/r2f_tests_occ24.log
python -c "
import json
for f in (16,20,24):
    try:
        d=json.loads(open('gpurun_out/r2f_bench_occ%d.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['same_final_positions_as_value'])
    except Exception as e: print(f, 'failed', e)
"
TL_LO=0 TL_HI=500 timeout 200 python tools/wave_timeline.py 1 -1 > gpurun_out/r2f_wave_timeline.txt 2>&1; head -2 gpurun_out/r2f_wave_timeline.txt
