set -x
mkdir -p gpurun_out/r2d
python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/r2d/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest_tc.log
python -m pytest tests -m gpu -q > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest.log
for a in 0 16; do NB_TC_ABLATE=$a timeout 300 python scripts/abl_probe.py >> gpurun_out/r2d/abl.jsonl 2>> gpurun_out/r2d/abl.err; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d/bench1.json 2> gpurun_out/r2d/bench1.err
