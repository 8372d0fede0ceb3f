set -x
mkdir -p gpurun_out/r2ncu gpurun_out/r2final
python -m pytest tests -m gpu -q > gpurun_out/r2final/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2final/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final/bench1.json 2> gpurun_out/r2final/bench1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2final/bench_ref.json 2> gpurun_out/r2final/bench_ref.err
NB_TC_ABLATE=0 timeout 300 python scripts/abl_probe.py > gpurun_out/r2final/abl.jsonl 2> gpurun_out/r2final/abl.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-render"
$CMD > gpurun_out/r2ncu/plain.json 2> gpurun_out/r2ncu/plain.err; echo "plain rc=$?" >> gpurun_out/r2ncu/plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2ncu/launches.csv $CMD > gpurun_out/r2ncu/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 18 -c 6 -o gpurun_out/r2ncu/prof -f $CMD > gpurun_out/r2ncu/ncu2.log 2>&1
