# usage: bash tools/gpu_prof_ingest.sh <tag>   ncu --set full of the last k_ingest_bulk launch of the C1 and C4 shapes
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
for cfg in C1 C4; do
  PROF_CONFIG=$cfg timeout 300 python tools/prof_ingest_configs.py > gpurun_out/ingest_${cfg}_$tag.jsonl 2>&1; cat gpurun_out/ingest_${cfg}_$tag.jsonl
  PROF_CONFIG=$cfg timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_bulk -s 2 -c 1 -o gpurun_out/prof_ingest_${cfg}_$tag -f python tools/prof_ingest_configs.py > gpurun_out/ncu_ingest_${cfg}_$tag.log 2>&1
  tail -2 gpurun_out/ncu_ingest_${cfg}_$tag.log
done
