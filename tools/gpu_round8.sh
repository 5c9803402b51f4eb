set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
g++ -O2 -std=c++17 -pthread -I high-fidelity-pointcloud-fusion_b200/csrc tools/pack_bench.cpp high-fidelity-pointcloud-fusion_b200/csrc/pcf_pack.cpp -o /tmp/pack_bench
for nt in 0 1 0 1; do echo -n "NT=$nt "; PCF_PACK_NT=$nt /tmp/pack_bench 12 96 | sort -k3 -n -r | head -1; done
timeout 600 python -m pytest tests/test_staging_gpu.py -x -q 2>&1 | tail -3
for nt in 0 1 0 1; do PCF_PACK_NT=$nt timeout 600 python bench.py --steps 8 --warmup 3 --no-c3 --no-cpu > gpurun_out/bench_nt${nt}_$tag.json 2> gpurun_out/bench_nt${nt}_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_nt${nt}_$tag.json"))
print("NT", $nt, "e2e", d["e2e"]["value"]/1e9, "pcie GB/s", d["e2e_roofline"]["achieved"], "whole_path ms", d["whole_path"]["ms"], "threads", d["e2e_roofline"]["stage_threads"])
PY
done
