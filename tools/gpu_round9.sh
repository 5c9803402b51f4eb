set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -6 gpurun_out/pytest_$tag.log
timeout 600 python tools/score_ab.py > gpurun_out/score_ab_$tag.jsonl 2> gpurun_out/score_ab_$tag.err; tail -2 gpurun_out/score_ab_$tag.jsonl; grep -E "k_sort|k_gather|k_score_coop" gpurun_out/score_ab_$tag.err | tail -12
timeout 600 python tools/trace_c3.py 1000 2> gpurun_out/trace_c3_$tag.err | tail -3
