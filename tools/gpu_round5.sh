set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_staging_gpu.py tests/test_parity_gpu.py -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -8 gpurun_out/pytest_$tag.log
timeout 600 python tools/score_ab.py > gpurun_out/score_ab_$tag.jsonl 2> gpurun_out/score_ab_$tag.err; tail -3 gpurun_out/score_ab_$tag.jsonl; grep -E "k_score" gpurun_out/score_ab_$tag.err | tail -3
for th in 8 10 12; do PCF_STAGE_THREADS=$th timeout 600 python bench.py --steps 5 --warmup 3 --no-c3 --no-cpu > gpurun_out/bench_th${th}_$tag.json 2> gpurun_out/bench_th${th}_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_th${th}_$tag.json"))
print("threads", $th, "e2e", d["e2e"]["value"]/1e9, "pcie GB/s", d["e2e_roofline"]["achieved"], "whole_path ms", d["whole_path"]["ms"], "value", d["value"]/1e9, "process", d["process_ms"])
PY
done
