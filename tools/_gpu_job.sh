set -x
mkdir -p gpurun_out/r2i
python -m pytest tests/test_arena_gpu.py -q 2>&1 | tail -15 > gpurun_out/r2i/pytest_arena.log
tail -3 gpurun_out/r2i/pytest_arena.log
python tools/steady_state.py --games 4096 --moves 12 --mode collapsed --max-game-len 16 --out gpurun_out/r2i/steady_collapsed_len16.json > gpurun_out/r2i/steady1.log 2>&1
python tools/steady_state.py --games 4096 --moves 12 --mode collapsed --out gpurun_out/r2i/steady_collapsed.json > gpurun_out/r2i/steady2.log 2>&1
python tools/steady_state.py --games 4096 --moves 10 --mode as_shipped --out gpurun_out/r2i/steady_as_shipped.json > gpurun_out/r2i/steady3.log 2>&1
tail -2 gpurun_out/r2i/steady*.log
