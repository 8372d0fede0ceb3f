set -x
mkdir -p gpurun_out/r2s
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2s/bench8.json 2> gpurun_out/r2s/bench8.err; echo "rc=$?" >> gpurun_out/r2s/bench8.err
NB_DP_MODE=joint timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 8 --steps 20 --warmup 5 --no-render > gpurun_out/r2s/bench8_joint.json 2> gpurun_out/r2s/bench8_joint.err; echo "rc=$?" >> gpurun_out/r2s/bench8_joint.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 4 --steps 20 --warmup 5 --no-render > gpurun_out/r2s/bench4.json 2> gpurun_out/r2s/bench4.err; echo "rc=$?" >> gpurun_out/r2s/bench4.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-render > gpurun_out/r2s/bench1.json 2> gpurun_out/r2s/bench1.err
