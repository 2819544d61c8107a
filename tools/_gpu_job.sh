mkdir -p gpurun_out/r2t
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2t/pytest.log
true
