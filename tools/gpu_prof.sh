# usage: bash tools/gpu_prof.sh <tag>  (on the GPU box under gpurun)  -> gpurun_out/launches_<tag>.csv + gpurun_out/prof_full_<tag>.ncu-rep
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
export PROF_REPS=2
timeout 200 python tools/prof_replay.py > gpurun_out/prof_plain_$tag.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/prof_replay.py > gpurun_out/ncu_list_$tag.log 2>&1
tail -3 gpurun_out/prof_plain_$tag.log
# full capture of the second repetition (13 matching launches per repetition: ingest, normals, 4 hist, 4 scatter, gather, score_work, score: skip the first)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ingest_bulk|k_score|k_sort_scatter|k_sort_hist|k_gather_points|k_normals|k_segment_heads" -s 13 -c 13 -o gpurun_out/prof_full_$tag -f python tools/prof_replay.py > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/ncu_full_$tag.log
ls -la gpurun_out | tail -5
