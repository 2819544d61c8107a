set -x
mkdir -p gpurun_out/r2n
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2n/pytest.log
tail -3 gpurun_out/r2n/pytest.log
python tools/steady_state.py --games 4096 --moves 10 --mode as_shipped --out gpurun_out/r2n/steady_as_shipped.json > gpurun_out/r2n/steady.log 2>&1
python bench.py --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/r2n/bench.json 2> gpurun_out/r2n/bench.err
true
