# usage: bash tools/gpu_round.sh <tag>   (runs on the GPU box under gpurun)
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
(nproc; free -g; nvidia-smi -L) > gpurun_out/box_$tag.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_$tag.log 2>&1; tail -30 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
timeout 600 python tools/score_ab.py > gpurun_out/score_ab_$tag.jsonl 2> gpurun_out/score_ab_$tag.err; cat gpurun_out/score_ab_$tag.jsonl; tail -60 gpurun_out/score_ab_$tag.err
