set -x
mkdir -p gpurun_out/r2j
python -m pytest tests/test_selfplay_gpu.py tests/test_mcts_gpu.py tests/test_mcts_stochastic_gpu.py tests/test_gameloop_gpu.py -q 2>&1 | tail -15 > gpurun_out/r2j/pytest.log
tail -3 gpurun_out/r2j/pytest.log
python tools/steady_state.py --games 4096 --moves 20 --mode collapsed --max-game-len 8 --out gpurun_out/r2j/steady_collapsed_len8.json > gpurun_out/r2j/steady1.log 2>&1
tail -c 300 gpurun_out/r2j/steady1.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/r2j/bench.json 2> gpurun_out/r2j/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'search_select|search_expand' -c 4 -o gpurun_out/r2j/tree python bench.py --steps 12 --warmup 3 --no-extras --cpu-seconds 1 > gpurun_out/r2j/ncu_tree.log 2>&1
true
