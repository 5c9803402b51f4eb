# usage: bash tools/gpu_multi.sh <N> <tag>   (under gpurun --gpus N)
set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}; tag=${2:-r}
mkdir -p gpurun_out
nproc > gpurun_out/box_multi_$tag.txt; free -g >> gpurun_out/box_multi_$tag.txt; nvidia-smi topo -m >> gpurun_out/box_multi_$tag.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
cat gpurun_out/bench_n${N}_$tag.json; tail -20 gpurun_out/bench_n${N}_$tag.err
