# single-GPU evidence run: box info, prefetch sweep of the pack loop, smoke, all GPU tests, bench (default + driver-style), launch list + full ncu capture
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
(nproc; free -g | head -2; lscpu | grep -E "Model name|L3") > gpurun_out/box_$tag.txt 2>&1
for d in 1024 2048 4096 8192; do g++ -O2 -std=c++17 -pthread -DPCF_PREFETCH_AHEAD=$d -I high-fidelity-pointcloud-fusion_b200/csrc tools/pack_bench.cpp high-fidelity-pointcloud-fusion_b200/csrc/pcf_pack.cpp -o /tmp/pack_bench_$d; echo -n "prefetch $d: " >> gpurun_out/box_$tag.txt; /tmp/pack_bench_$d 12 96 | sort -k3 -n -r | head -1 >> gpurun_out/box_$tag.txt; done
cat gpurun_out/box_$tag.txt
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_$tag.log 2>&1; tail -14 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; cat gpurun_out/bench_ref_$tag.json
bash tools/gpu_prof.sh $tag
