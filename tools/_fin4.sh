mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_gputests.log
grep -E "passed|failed|^FAILED|^E  |rc=" gpurun_out/r2s_gputests.log | head -6
timeout 600 python tools/make_inst_table.py --plies 5,9,13,17,21,24 > gpurun_out/r2s_inst.log 2>&1; cp profiles/inst_table* gpurun_out/
tail -13 gpurun_out/r2s_inst.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['mcts_sims_per_s'], d['e2e']['ms_per_step'], d['e2e']['same_final_positions_as_value'], d['roofline']['frac'], d['roofline']['table']['stale'], d['az_mcts']['sims_per_s'])
"
