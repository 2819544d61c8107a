set -x
mkdir -p gpurun_out/r2d
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2d/pytest.log
tail -3 gpurun_out/r2d/pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err
tail -c 600 gpurun_out/r2d/bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2d/bench_ref.json 2> gpurun_out/r2d/bench_ref.err
