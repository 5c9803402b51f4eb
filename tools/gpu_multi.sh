# usage: bash tools/gpu_multi.sh <N> <tag>   bench at N GPUs + C4 interleaved sharded
set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}; tag=${2:-r}
mkdir -p gpurun_out
nproc > gpurun_out/box_multi_$tag.txt; free -g >> gpurun_out/box_multi_$tag.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
cat gpurun_out/bench_n${N}_$tag.json; tail -5 gpurun_out/bench_n${N}_$tag.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_c4.py 50 10 > gpurun_out/c4_n${N}_$tag.json 2> gpurun_out/c4_n${N}_$tag.err
cat gpurun_out/c4_n${N}_$tag.json; tail -5 gpurun_out/c4_n${N}_$tag.err
