# usage: bash tools/gpu_prof.sh   (on the GPU box under gpurun)  -> gpurun_out/launches.csv + gpurun_out/prof_full.ncu-rep
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PROF_REPS=3
timeout 200 python tools/prof_replay.py > gpurun_out/prof_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_replay.py > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/prof_plain.log
# full capture of the third repetition (14 matching launches per repetition: skip the first two)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ingest|k_score|k_sort_scatter|k_normals|k_gather_points|k_cells_to_bits|k_sort_hist|k_segment_heads" -s 28 -c 14 -o gpurun_out/prof_full -f python tools/prof_replay.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -5
