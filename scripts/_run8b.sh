mkdir -p gpurun_out/r2h
python -m pytest tests -m gpu -q > gpurun_out/r2h/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h/pytest.log
