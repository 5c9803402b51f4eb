# usage: bash tools/gpu_multi_rawlanes.sh <N> <tag>   e2e A/B of packers x raw lanes at N GPUs (no C3 leg)
set -x
cd $GRAFT_REPO_ROOT
N=${1:-8}; tag=${2:-r}
mkdir -p gpurun_out
for cfg in "1 8" "2 8" "1 12"; do
set -- $cfg; th=$1; lanes=$2
PCF_STAGE_THREADS=$th PCF_RAW_LANES=$lanes timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-c3 2>gpurun_out/bench_n${N}_t${th}_l${lanes}_$tag.err | grep "^{" > gpurun_out/bench_n${N}_t${th}_l${lanes}_$tag.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${N}_t${th}_l${lanes}_$tag.json"))
print("N", $N, "threads", $th, "lanes", $lanes, "e2e", d["e2e"]["value"]/1e9, "h2d/step", d["e2e"]["h2d_bytes_per_step"], "whole_path ms", d["whole_path"]["ms"])
PY
done
