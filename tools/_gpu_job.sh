set -x
mkdir -p gpurun_out/r2o
python -m pytest tests/test_selfplay_gpu.py tests/test_mcts_stochastic_gpu.py -q 2>&1 | tail -5 > gpurun_out/r2o/pytest.log
tail -2 gpurun_out/r2o/pytest.log
python tools/steady_state.py --games 4096 --moves 10 --mode as_shipped --out gpurun_out/r2o/steady_as_shipped.json > gpurun_out/r2o/steady.log 2>&1
tail -c 400 gpurun_out/r2o/steady.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/r2o/bench.json 2> gpurun_out/r2o/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_pair_kernel' --launch-skip 6 -c 2 -o gpurun_out/r2o/conv_pair python bench.py --steps 1 --warmup 1 --no-extras --cpu-seconds 1 > gpurun_out/r2o/ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|se_tail|value_tail|encode_mask_planes' --launch-skip 14 -c 8 -o gpurun_out/r2o/gemm python bench.py --steps 1 --warmup 1 --no-extras --cpu-seconds 1 > gpurun_out/r2o/ncu_gemm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2o/launches.csv python bench.py --steps 2 --warmup 1 --no-extras --cpu-seconds 1 > gpurun_out/r2o/ncu_list.log 2>&1
true
