# usage: bash tools/gpu_multi.sh N   (under gpurun --gpus N)
set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; grep -E "byte-identical|MULTI_GPU_CHECK|Error|error" gpurun_out/multi_check_$N.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 2500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
