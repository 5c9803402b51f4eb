# usage: bash tools/gpu_prof_one.sh <kernel-regex> <tag> [skip]   full ncu capture of ONE launch of the matching kernel in the C2 replay
set -x
cd $GRAFT_REPO_ROOT
k=$1; tag=${2:-r}; skip=${3:-1}
mkdir -p gpurun_out
export PROF_REPS=2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -o gpurun_out/prof_${tag} -f python tools/prof_replay.py > gpurun_out/ncu_${tag}.log 2>&1
tail -3 gpurun_out/ncu_${tag}.log
