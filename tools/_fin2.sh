mkdir -p gpurun_out
timeout 300 python tools/gen_bench_positions.py > gpurun_out/r2n_positions.log 2>&1; cp oracle/bench_positions.npz gpurun_out/; tail -1 gpurun_out/r2n_positions.log
timeout 700 python tools/make_inst_table.py --plies 5,9,13,17,21,24 > gpurun_out/r2n_inst.log 2>&1; cp profiles/inst_table* gpurun_out/
tail -14 gpurun_out/r2n_inst.log
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2n_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['table'])
"
