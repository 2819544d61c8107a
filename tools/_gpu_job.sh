set -x
mkdir -p gpurun_out/r2b
python -m pytest tests/test_mcts_stochastic_gpu.py tests/test_mcts_gpu.py tests/test_selfplay_gpu.py -q 2>&1 | tail -60 > gpurun_out/r2b/pytest_stoch.log
tail -5 gpurun_out/r2b/pytest_stoch.log
