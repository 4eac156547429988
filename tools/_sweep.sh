run() { echo "== $*"; env $1 timeout 400 python bench.py --no-kernels --no-az --no-cpu-baseline $2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.3g'%d['value'], round(d['ms_per_step'],1), '%.3g'%d['e2e']['value'], round(d['roofline']['avg_main_stream_ms'],2), round(d['roofline']['avg_deferred_stuck_pass_ms'],2))"; }
run "QZ_MAIN_PRIO=0" "--streams 2 --defer 4"
run "QZ_MAIN_PRIO=-1" "--streams 2 --defer 4"
run "QZ_MAIN_PRIO=-1" "--streams 1 --defer 4"
run "QZ_MAIN_PRIO=-1" "--streams 2 --defer 8"
run "QZ_MAIN_PRIO=-1 CUDA_DEVICE_MAX_CONNECTIONS=32" "--streams 4 --defer 8"
