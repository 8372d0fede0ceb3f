set -x
mkdir -p gpurun_out/r2s
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2972$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2s/bench$n.json 2> gpurun_out/r2s/bench$n.err; echo "rc=$?" >> gpurun_out/r2s/bench$n.err
done
NB_DP_MODE=joint timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29729 bench.py --gpus 8 --steps 20 --warmup 5 --no-render > gpurun_out/r2s/bench8_joint.json 2> gpurun_out/r2s/bench8_joint.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2s/bench1.json 2> gpurun_out/r2s/bench1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29730 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r2s/bench_ref8.json 2> gpurun_out/r2s/bench_ref8.err
python -m pytest tests/test_gpu_extra.py -q -k "two_gpu" > gpurun_out/r2s/pytest_2gpu.log 2>&1
