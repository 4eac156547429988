set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mcts_gpu.py tests/test_bench_parity.py -q -x -s > gpurun_out/r2d_mctstests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_mctstests.log
grep -E "passed|failed|^FAILED|^E  |TV" gpurun_out/r2d_mctstests.log | head -30
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2d_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_gputests.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2d_gputests.log | head -30
timeout 300 python tools/debug_e2e.py 4096 20 > gpurun_out/r2d_debug_e2e.log 2>&1; tail -12 gpurun_out/r2d_debug_e2e.log
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2d_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'])
"
timeout 900 python tools/make_inst_table.py --plies 26 > gpurun_out/r2d_inst.log 2>&1; cp profiles/inst_table* gpurun_out/
tail -22 gpurun_out/r2d_inst.log
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qz_rollout_stuck" --launch-skip 6 -c 2 -f -o gpurun_out/prof_stuck_r2d python tools/prof_wave.py > gpurun_out/prof_stuck_r2d.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qz_mcts_select|qz_mcts_extend|qz_mcts_expand|qz_rollout_wall" --launch-skip 32 -c 4 -f -o gpurun_out/prof_tree_r2d python tools/prof_wave.py > gpurun_out/prof_tree_r2d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
