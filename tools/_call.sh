timeout 300 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -3
B="timeout 150 python bench.py --no-variants --no-cpu-baseline --steps 50 --warmup 10"
P='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],d["ms_per_step"],d["ms_per_step_median"],d["gpu_launches_per_step"])'
$B 2>/dev/null | python -c "$P" latewait
B200SEG_LIB=$PWD/gpurun_in/libb200seg_waitfirst.so $B 2>/dev/null | python -c "$P" waitfirst
$B 2>/dev/null | python -c "$P" latewait
B200SEG_LIB=$PWD/gpurun_in/libb200seg_waitfirst.so $B 2>/dev/null | python -c "$P" waitfirst
