set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
export PROF_REPS=2 PROF_FRAMES=12
for tool in memcheck initcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/prof_replay.py > gpurun_out/sanitize_${tool}_$tag.log 2>&1
  grep -E "variant|ERROR SUMMARY|Uninitialized|Invalid|hazard|Race" gpurun_out/sanitize_${tool}_$tag.log | head -12
done
