set -x
mkdir -p gpurun_out/r2g
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2g/pytest.log
tail -3 gpurun_out/r2g/pytest.log
python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2g/bench_a.json 2> gpurun_out/r2g/bench_a.err
M0_SE_NOSPLIT=1 python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2g/bench_nosplit.json 2> gpurun_out/r2g/bench_nosplit.err
python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r2g/bench_b.json 2> gpurun_out/r2g/bench_b.err
# compute-sanitizer: memcheck + racecheck over the tensor-core kernels (single-CTA GEMM, CTA-pair convolution with its fused epilogues,
# attention), the tree kernels and the game loop
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_tc_gemm_gpu.py -q -x -k "2-64-64 or 2-320-320 or 6-128-160 or 4-128-48 or 4-128-320" > gpurun_out/r2g/sanitizer_memcheck_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2g/sanitizer_memcheck_tc.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g/sanitizer_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2g/sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gameloop_gpu.py tests/test_mcts_gpu.py::test_api_surface "tests/test_mcts_stochastic_gpu.py::test_device_generator_statistics" -q -x > gpurun_out/r2g/sanitizer_memcheck_tree.log 2>&1; echo "rc=$?" >> gpurun_out/r2g/sanitizer_memcheck_tree.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g/sanitizer_racecheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2g/sanitizer_racecheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gameloop_gpu.py::test_start_budget_plays_every_started_game_to_its_end tests/test_mcts_gpu.py::test_api_surface -q -x > gpurun_out/r2g/sanitizer_racecheck_tree.log 2>&1; echo "rc=$?" >> gpurun_out/r2g/sanitizer_racecheck_tree.log
for f in gpurun_out/r2g/sanitizer_*.log; do echo == $f; tail -4 $f; done
