set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mcts_gpu.py -q -x > gpurun_out/r2c_mctstests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_mctstests.log
tail -25 gpurun_out/r2c_mctstests.log
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_gputests.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2c_gputests.log | head -30
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2c_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['az_mcts'])
"
timeout 900 python tools/make_inst_table.py --plies 26 > gpurun_out/r2c_inst.log 2>&1; cp profiles/inst_table* gpurun_out/
tail -25 gpurun_out/r2c_inst.log
