set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for b in 1 0 1 0; do PCF_INGEST_BITS=$b timeout 600 python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu > gpurun_out/bench_bits${b}.json 2> gpurun_out/bench_bits${b}.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_bits${b}.json"))
print("BITS", $b, "value", d["value"]/1e9, "frac", d["roofline"]["frac"], "process", d["process_ms"], d["process_detail"]["update_ms"], "voxels", d["process_detail"]["voxels"])
PY
done
