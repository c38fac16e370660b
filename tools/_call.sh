timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "gn_bwd_reduce_large or groupnorm" 2>&1 | tail -6
B="timeout 150 python bench.py --no-variants --no-cpu-baseline --steps 50 --warmup 10"
P='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],d["ms_per_step"],d["ms_per_step_median"],d["gpu_launches_per_step"])'
$B 2>/dev/null | python -c "$P" default
B200SEG_GN_REDUCE_DEEP=0 $B 2>/dev/null | python -c "$P" deep0
$B 2>/dev/null | python -c "$P" default_again
timeout 100 python tools/microbench_ops.py --only gn 2>&1 | tail -25
