# usage: bash tools/gpu_ab.sh "VAR=val VAR=val" "..." ...   (A/B of library variants; each argument is one env set, "-" = default)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1; tail -5 gpurun_out/pytest_ab.log
i=0
for v in "$@"; do
  i=$((i+1))
  if [ "$v" = "-" ]; then e=""; else e="$v"; fi
  env $e timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_ab_$i.json 2> gpurun_out/bench_ab_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ab_$i.json"))
    print("AB [$v]", "value %.1f Gpts/s" % (d["value"]/1e9), "frac %.3f" % d["roofline"]["frac"], "launch_ms %.4f" % d["roofline"]["launch_ms"], "e2e %.2f" % (d["e2e"]["value"]/1e9), "process_ms %.2f" % d["process_ms"])
except Exception as ex:
    print("AB [$v] failed", ex)
PY
done
