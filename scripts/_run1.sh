set -x
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
for a in 0 16 64 192 80; do NB_TC_ABLATE=$a timeout 300 python scripts/abl_probe.py >> gpurun_out/r2a/abl.jsonl 2>> gpurun_out/r2a/abl.err; done
timeout 900 python scripts/ref_probe.py > gpurun_out/r2a/ref_probe.log 2>&1; echo "ref_probe rc=$?" >> gpurun_out/r2a/ref_probe.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
