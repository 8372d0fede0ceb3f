set -x
mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e/pytest.log
timeout 600 python scripts/train_demo.py --steps 3000 --fp32-steps 300 > gpurun_out/r2e/train_demo.jsonl 2> gpurun_out/r2e/train_demo.err
timeout 300 python scripts/fuzz_shapes.py > gpurun_out/r2e/fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r2e/fuzz.log
timeout 600 python scripts/sweep.py > gpurun_out/r2e/sweep.jsonl 2> gpurun_out/r2e/sweep.err
timeout 600 python bench.py --workload llff --steps 20 --warmup 5 --no-cpu > gpurun_out/r2e/bench_llff.json 2> gpurun_out/r2e/bench_llff.err
timeout 900 python scripts/ref_probe.py --no-time > gpurun_out/r2e/ref_probe.log 2>&1
