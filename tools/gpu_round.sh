# usage: bash tools/gpu_round.sh <tag>   (runs on the GPU box under gpurun)
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -15 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
