mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2l_pytest.log
