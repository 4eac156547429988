set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/r2j_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_gputests.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2j_gputests.log | head -20
timeout 300 python tools/gen_bench_positions.py > gpurun_out/r2j_positions.log 2>&1; cp oracle/bench_positions.npz gpurun_out/; tail -1 gpurun_out/r2j_positions.log
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2j_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'])
"
timeout 300 python bench.py --impl reference --steps 4 --warmup 5 --pyref-seconds 3 > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err
timeout 300 python bench.py --workload az --steps 3 --warmup 2 --no-kernels --no-parity --no-cpu-baseline --games-plies 0 > gpurun_out/r2j_bench_az.json 2> gpurun_out/r2j_bench_az.err; tail -2 gpurun_out/r2j_bench_az.err
TL_LO=0 TL_HI=500 timeout 200 python tools/wave_timeline.py 1 -1 > gpurun_out/r2j_wave_timeline.txt 2>&1; head -1 gpurun_out/r2j_wave_timeline.txt
timeout 700 python tools/make_inst_table.py --plies 5,9,13,17,21,24 > gpurun_out/r2j_inst.log 2>&1; cp profiles/inst_table* gpurun_out/
tail -16 gpurun_out/r2j_inst.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qz_rollout_stuck" --launch-skip 6 -c 2 -f -o gpurun_out/prof_stuck_r2j python tools/prof_wave.py > gpurun_out/prof_stuck_r2j.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qz_mcts_select|qz_mcts_extend|qz_mcts_expand|qz_rollout_wall|qz_legal_mask_flagged|qz_rollout_pawn" --launch-skip 70 -c 13 -f -o gpurun_out/prof_wave_r2j python tools/prof_wave.py > gpurun_out/prof_wave_r2j.log 2>&1
ls -la gpurun_out/*r2j*.ncu-rep
