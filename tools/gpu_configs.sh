# usage: bash tools/gpu_configs.sh <tag>   C1..C5 timings (warm second pass) + full-size oracle comparisons of C3 / C4 (restatement) and C3 (reference header)
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
rm -f gpurun_out/configs.jsonl
timeout 1500 python tools/run_configs.py C1 C2 C3 C4 C5 > gpurun_out/configs_$tag.jsonl 2> gpurun_out/configs_$tag.err; cat gpurun_out/configs_$tag.jsonl | cut -c1-700
timeout 1500 python tools/run_configs.py C3 C4 --oracle > gpurun_out/configs_oracle_$tag.jsonl 2> gpurun_out/configs_oracle_$tag.err; cat gpurun_out/configs_oracle_$tag.jsonl | cut -c1-1500
timeout 1800 python tools/run_configs.py C3 --oracle --oracle-kind ref_ordered > gpurun_out/configs_ref_$tag.jsonl 2> gpurun_out/configs_ref_$tag.err; cat gpurun_out/configs_ref_$tag.jsonl | cut -c1-1500; tail -3 gpurun_out/configs_ref_$tag.err
