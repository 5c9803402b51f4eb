set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
timeout 600 python tools/score_ab.py > gpurun_out/score_ab_$tag.jsonl 2> gpurun_out/score_ab_$tag.err; tail -2 gpurun_out/score_ab_$tag.jsonl; grep -E "k_sort|k_gather|k_score_coop" gpurun_out/score_ab_$tag.err | tail -12
timeout 600 python tools/trace_c3.py 1000 2> gpurun_out/trace_c3_$tag.err | tail -3
sed -n "/---- trace ----/,\$p" gpurun_out/trace_c3_$tag.err | grep -E "k_sort|k_gather|k_score_coop"
