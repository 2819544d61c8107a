set -x
mkdir -p gpurun_out/r2q
python -m pytest tests/test_nn_gpu.py tests/test_selfplay_gpu.py -q 2>&1 | tail -5 > gpurun_out/r2q/pytest.log
tail -2 gpurun_out/r2q/pytest.log
python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2q/bench.json 2> gpurun_out/r2q/bench.err
M0_SE_TAIL=0 python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2q/bench_nosetail.json 2> gpurun_out/r2q/bench_nosetail.err
python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2q/bench2.json 2> gpurun_out/r2q/bench2.err
true
