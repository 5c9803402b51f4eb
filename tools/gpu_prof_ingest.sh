# usage: bash tools/gpu_prof_ingest.sh "<PROF_VARIANTS>" [A/B env sets...]  -- A/B bench lines, then ncu --set full of the ingest kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
variants="$1"; shift
if [ $# -gt 0 ]; then bash tools/gpu_ab.sh "$@"; fi
export PROF_VARIANTS="$variants" PROF_REPS=2
python tools/prof_replay.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_ingest" -o gpurun_out/prof_ingest -f python tools/prof_replay.py > gpurun_out/ncu_ingest.log 2>&1
tail -3 gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_ingest.log
