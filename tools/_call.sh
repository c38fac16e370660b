set -o pipefail
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py 2>&1 | tail -8 ) > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
( timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -5 ) > gpurun_out/final_smoke.txt; cat gpurun_out/final_smoke.txt
B200SEG_BENCH_TABLE=gpurun_out/final_per_op_table.txt timeout 400 python bench.py 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench_n1.json
python -c "
import json;d=json.load(open('gpurun_out/final_bench_n1.json'));print('bench',d['ms_per_step'],d['ms_per_step_median'],d['e2e'],d['roofline']['frac'],d.get('variants'))"
timeout 200 python bench.py --workload unet2d512 --no-variants --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_unet2d512_n1.json
python -c "
import json;d=json.load(open('gpurun_out/final_bench_unet2d512_n1.json'));print('unet2d',d['ms_per_step'],d['value'])"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
timeout 400 ncu --metrics $M --clock-control none -f -o gpurun_out/final_ncu_ops python tools/microbench_ops.py --eager --only conv,convbwd,wgrad,gn,stem > gpurun_out/final_ncu_ops.log 2>&1; tail -2 gpurun_out/final_ncu_ops.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv python tools/profile_step.py > gpurun_out/final_launches.log 2>&1; tail -2 gpurun_out/final_launches.log
ls -la gpurun_out | tail -12
