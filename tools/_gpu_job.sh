set -x
mkdir -p gpurun_out/r2e
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --section ComputeWorkloadAnalysis"
timeout 900 ncu $SEC --clock-control none --import-source on -k regex:'attention_tc|ln_res_gn' -c 4 -o gpurun_out/r2e/attn python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2e/ncu_attn.log 2>&1
timeout 900 ncu $SEC --clock-control none -k regex:'gemm_tc_kernel' -c 60 -o gpurun_out/r2e/gemm python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2e/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'search_select|search_expand' -c 4 -o gpurun_out/r2e/tree python bench.py --steps 12 --warmup 3 --no-extras > gpurun_out/r2e/ncu_tree.log 2>&1
ls -la gpurun_out/r2e
