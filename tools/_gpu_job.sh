set -x
mkdir -p gpurun_out/r2c
python -m pytest tests/test_gameloop_gpu.py tests/test_selfplay_gpu.py tests/test_nn_gpu.py -q 2>&1 | tail -60 > gpurun_out/r2c/pytest.log
tail -5 gpurun_out/r2c/pytest.log
