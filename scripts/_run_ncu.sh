set -x
mkdir -p gpurun_out/r2ncu
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-render"
$CMD > gpurun_out/r2ncu/plain.json 2> gpurun_out/r2ncu/plain.err; echo "plain rc=$?" >> gpurun_out/r2ncu/plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2ncu/launches.csv $CMD > gpurun_out/r2ncu/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 18 -c 6 -o gpurun_out/r2ncu/prof -f $CMD > gpurun_out/r2ncu/ncu2.log 2>&1
ls -la gpurun_out/r2ncu
