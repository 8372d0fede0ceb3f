set -x
mkdir -p gpurun_out/r2b
python -m pytest tests -m gpu -q > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2b/smoke.log
NB_TC_ABLATE=0 timeout 300 python scripts/abl_probe.py >> gpurun_out/r2b/abl.jsonl 2>> gpurun_out/r2b/abl.err
timeout 900 python scripts/ref_probe.py --no-time > gpurun_out/r2b/ref_probe.log 2>&1; echo "ref_probe rc=$?" >> gpurun_out/r2b/ref_probe.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b/bench_ref.json 2> gpurun_out/r2b/bench_ref.err
