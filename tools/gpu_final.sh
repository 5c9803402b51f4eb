# usage: bash tools/gpu_final.sh <tag>   full single-GPU round: smoke, GPU tests, bench (+ reference arm), configs C1..C5
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-z}
mkdir -p gpurun_out
rm -f gpurun_out/configs.jsonl
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-900 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; cut -c1-400 gpurun_out/bench_ref_$tag.json
timeout 150 python -u tools/run_configs.py C1 --oracle > gpurun_out/cfg_C1.log 2>&1
timeout 200 python -u tools/run_configs.py C2 --oracle > gpurun_out/cfg_C2.log 2>&1
for c in C3 C4 C5; do timeout 400 python -u tools/run_configs.py $c > gpurun_out/cfg_$c.log 2>&1; done
grep -c oracle_bit_exact gpurun_out/configs.jsonl; wc -l gpurun_out/configs.jsonl
