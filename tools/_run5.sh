set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mcts_gpu.py tests/test_bench_parity.py -q -x -s > gpurun_out/r2e_mctstests.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_mctstests.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2e_mctstests.log | head -20
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/r2e_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_gputests.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2e_gputests.log | head -20
timeout 300 python tools/debug_e2e.py 4096 20 > gpurun_out/r2e_debug_e2e.log 2>&1; tail -4 gpurun_out/r2e_debug_e2e.log
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2e_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'])
"
TL_LO=0 TL_HI=500 timeout 200 python tools/wave_timeline.py 1 -1 > gpurun_out/r2e_wave_timeline.txt 2>&1; head -3 gpurun_out/r2e_wave_timeline.txt
timeout 700 python tools/make_inst_table.py --plies 5,9,13,17,21,24 > gpurun_out/r2e_inst.log 2>&1; cp profiles/inst_table* gpurun_out/
tail -16 gpurun_out/r2e_inst.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qz_rollout_stuck" --launch-skip 6 -c 2 -f -o gpurun_out/prof_stuck_r2e python tools/prof_wave.py > gpurun_out/prof_stuck_r2e.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qz_mcts_select|qz_mcts_extend|qz_mcts_expand|qz_rollout_wall|qz_legal_mask_flagged" --launch-skip 40 -c 5 -f -o gpurun_out/prof_tree_r2e python tools/prof_wave.py > gpurun_out/prof_tree_r2e.log 2>&1
ls -la gpurun_out/*r2e*.ncu-rep
