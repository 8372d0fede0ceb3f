# ncu --set full of the HBM-bound helper kernels (ray-gen, sampling, encoding, compositing) at sizes larger than L2; only the raw-page CSV travels back
mkdir -p gpurun_out/r2small
python scripts/kernel_bench.py --what hbm > gpurun_out/r2small/plain.jsonl 2> gpurun_out/r2small/plain.err; echo "rc=$?" >> gpurun_out/r2small/plain.err
ncu --set full --clock-control none -k regex:'composite_|sample_pdf|stratified|raygen' -c 40 -o /tmp/prof_small -f python scripts/kernel_bench.py --what hbm > gpurun_out/r2small/ncu.log 2>&1
ncu -i /tmp/prof_small.ncu-rep --page raw --csv > gpurun_out/r2small/raw.csv 2>/dev/null
ls -la /tmp/prof_small.ncu-rep gpurun_out/r2small/
