set -x
mkdir -p gpurun_out/r2c
python -m pytest tests -m gpu -q > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c/bench1.json 2> gpurun_out/r2c/bench1.err
for mode in joint peer; do
NB_DP_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 --no-render > gpurun_out/r2c/bench2_$mode.json 2> gpurun_out/r2c/bench2_$mode.err
NB_DP_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 --no-render --no-graph > gpurun_out/r2c/bench2_${mode}_nograph.json 2> gpurun_out/r2c/bench2_${mode}_nograph.err
done
