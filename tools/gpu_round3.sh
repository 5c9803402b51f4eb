# usage: bash tools/gpu_round3.sh <tag>   quick: staging + parity tests, score A/B, bench (no C3 leg), e2e lanes A/B
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_staging_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -8 gpurun_out/pytest_$tag.log
timeout 600 python tools/score_ab.py > gpurun_out/score_ab_$tag.jsonl 2> gpurun_out/score_ab_$tag.err; cat gpurun_out/score_ab_$tag.jsonl; grep -E "k_score|k_gather|k_sort_scatter" gpurun_out/score_ab_$tag.err | tail -12
for lanes in 0 1 2 3; do PCF_RAW_LANES=$lanes timeout 600 python bench.py --steps 5 --warmup 3 --no-c3 --no-cpu > gpurun_out/bench_lanes${lanes}_$tag.json 2> gpurun_out/bench_lanes${lanes}_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_lanes${lanes}_$tag.json"))
print("lanes", $lanes, "e2e", d["e2e"]["value"]/1e9, "h2d/step", d["e2e"]["h2d_bytes_per_step"], "pcie GB/s", d["e2e_roofline"]["achieved"], "whole_path ms", d["whole_path"]["ms"], "value", d["value"]/1e9, "process", d["process_ms"])
PY
done
