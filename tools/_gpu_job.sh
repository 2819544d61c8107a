set -x
mkdir -p gpurun_out/r2l
python -m pytest tests/test_nn_gpu.py tests/test_tc_gemm_gpu.py tests/test_mcts_stochastic_gpu.py tests/test_selfplay_gpu.py tests/test_inference_gpu.py -q 2>&1 | tail -30 > gpurun_out/r2l/pytest.log
tail -3 gpurun_out/r2l/pytest.log
python tools/refinit_probe.py fp16 bf16 > gpurun_out/r2l/probe.jsonl 2> gpurun_out/r2l/probe.err
python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2l/bench_new.json 2> gpurun_out/r2l/bench_new.err
M0_TC_PROJ_F32=1 M0_TC_VALUE_TAIL=0 python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2l/bench_old.json 2> gpurun_out/r2l/bench_old.err
python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2l/bench_new2.json 2> gpurun_out/r2l/bench_new2.err
python tools/steady_state.py --games 4096 --moves 10 --mode as_shipped --out gpurun_out/r2l/steady_as_shipped.json > gpurun_out/r2l/steady.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2l/launches_as_shipped.csv python tools/steady_state.py --games 4096 --moves 1 --mode as_shipped --out gpurun_out/r2l/tmp.json > gpurun_out/r2l/ncu_as.log 2>&1
true
