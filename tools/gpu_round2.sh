# usage: bash tools/gpu_round2.sh <tag>   tests + bench + host pack bench + ncu
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
(nproc; lscpu | grep -E "Model name|Socket|NUMA|L3|Thread|Core"; lscpu | grep -o -E "avx512[a-z_0-9]*|avx2" | sort -u | tr '\n' ' ') > gpurun_out/box_$tag.txt 2>&1
g++ -O2 -std=c++17 -pthread -I high-fidelity-pointcloud-fusion_b200/csrc tools/pack_bench.cpp high-fidelity-pointcloud-fusion_b200/csrc/pcf_pack.cpp -o /tmp/pack_bench && for t in 1 8 16; do /tmp/pack_bench $t 64 | tail -1; done >> gpurun_out/box_$tag.txt 2>&1
cat gpurun_out/box_$tag.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -8 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
bash tools/gpu_prof.sh $tag
