set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_gpu.py -q -x -k "rollout or stuck" > gpurun_out/r2b_rollouttests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_rollouttests.log
tail -3 gpurun_out/r2b_rollouttests.log
timeout 900 python -m pytest tests/test_mcts_gpu.py -q -x > gpurun_out/r2b_mctstests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_mctstests.log
tail -30 gpurun_out/r2b_mctstests.log
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_gputests.log
tail -15 gpurun_out/r2b_gputests.log
timeout 300 python tools/debug_e2e.py > gpurun_out/r2b_debug_e2e.log 2>&1
cat gpurun_out/r2b_debug_e2e.log | tail -12
QZ_PAWN_SLICED=1 timeout 600 python bench.py --steps 8 --warmup 5 --no-az --no-kernels --no-parity --no-cpu-baseline --games-plies 0 > gpurun_out/r2b_bench_sliced.json 2> gpurun_out/r2b_bench_sliced.err
timeout 600 python bench.py --steps 8 --warmup 5 --no-az --no-kernels --no-parity --no-cpu-baseline --games-plies 0 > gpurun_out/r2b_bench_queue.json 2> gpurun_out/r2b_bench_queue.err
python -c "
import json
for f in ('sliced','queue'):
    try:
        d=json.loads(open('gpurun_out/r2b_bench_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e'])
    except Exception as e: print(f, 'failed', e)
"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b_bench.err
timeout 1000 python tools/make_inst_table.py --plies 26 > gpurun_out/r2b_inst.log 2>&1; cp profiles/inst_table* gpurun_out/
tail -20 gpurun_out/r2b_inst.log
