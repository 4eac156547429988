set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_env_gpu.py -q -x -k "rollout or stuck" > gpurun_out/r2a_stucktests.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_stucktests.log
tail -5 gpurun_out/r2a_stucktests.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2a_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_gputests.log
tail -15 gpurun_out/r2a_gputests.log
timeout 600 python tools/gen_bench_positions.py > gpurun_out/r2a_positions.log 2>&1; cp oracle/bench_positions.npz gpurun_out/
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 4 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
timeout 1800 python tools/make_inst_table.py --plies 26 > gpurun_out/r2a_inst.log 2>&1; cp profiles/inst_table* gpurun_out/
tail -20 gpurun_out/r2a_inst.log
