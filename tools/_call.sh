# One bounded GPU call (round 2, last GPU minutes): A/B of the scheduling options on one box -> pick -> validate the
# pick with the GPU suite, smoke and bench.  Everything lands in gpurun_out/ as it is produced.
set -o pipefail
mkdir -p gpurun_out
date +%s > gpurun_out/t0
( timeout 170 python tools/overlap_ab.py --masks 0,1,2,4,8,15,0 --bps 4,2 --choose gpurun_out/chosen_env.sh ) > gpurun_out/overlap_ab.jsonl 2> gpurun_out/overlap_ab.err
tail -4 gpurun_out/overlap_ab.jsonl; tail -3 gpurun_out/overlap_ab.err
[ -f gpurun_out/chosen_env.sh ] || echo "export B200SEG_OVERLAP=0" > gpurun_out/chosen_env.sh
. gpurun_out/chosen_env.sh
env | grep B200SEG_ > gpurun_out/chosen_env.txt; cat gpurun_out/chosen_env.txt
timeout 260 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_losses.py tests/test_gpu_ops.py tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/final_pytest.txt 2>&1; tail -6 gpurun_out/final_pytest.txt
( timeout 90 python __graft_entry__.py --smoke 2>&1 | tail -5 ) > gpurun_out/final_smoke.txt; cat gpurun_out/final_smoke.txt
B200SEG_BENCH_TABLE=gpurun_out/final_per_op_table.txt timeout 200 python bench.py 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench_n1.json
python -c "
import json;d=json.load(open('gpurun_out/final_bench_n1.json'));print('bench',d['ms_per_step'],d['ms_per_step_median'],d['e2e'],d['roofline']['frac'],d.get('variants'))"
( timeout 100 python tools/overlap_ab.py --workload unet2d --masks 0,$B200SEG_OVERLAP,0 ) > gpurun_out/overlap_ab_unet2d.jsonl 2>> gpurun_out/overlap_ab.err; tail -3 gpurun_out/overlap_ab_unet2d.jsonl
( timeout 100 python tools/overlap_ab.py --workload unet3d --masks 0,$B200SEG_OVERLAP,0 ) > gpurun_out/overlap_ab_unet3d.jsonl 2>> gpurun_out/overlap_ab.err; tail -3 gpurun_out/overlap_ab_unet3d.jsonl
echo elapsed $(( $(date +%s) - $(cat gpurun_out/t0) )) s
if [ "$B200SEG_OVERLAP" != "0" ]; then
  B200SEG_OVERLAP=0 timeout 150 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > gpurun_out/final_pytest_mask0.txt 2>&1; tail -4 gpurun_out/final_pytest_mask0.txt
fi
echo elapsed $(( $(date +%s) - $(cat gpurun_out/t0) )) s
