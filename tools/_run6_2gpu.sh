set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 500 python -m pytest tests/test_train_gpu.py -q -s -k "nccl" > gpurun_out/r2f_nccl_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_nccl_test.log
tail -5 gpurun_out/r2f_nccl_test.log
NCCL_DEBUG=INFO timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 5 --games-plies 40 --az-steps 2 > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo "bench rc=$?"
grep -E "NVLS|P2P|NET/|via" gpurun_out/r2f_bench_2gpu.err | head -8
tail -c 1500 gpurun_out/r2f_bench_2gpu.json
