set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/prof_replay.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_replay.py > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/prof_plain.log
# full capture of the third repetition's kernels (skip the first two reps' launches of the selected kernels)
python tools/prof_replay.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_ingest|k_score|k_sort_scatter|k_normals|k_gather_points|k_cells_to_bits|k_sort_hist|k_segment_heads" -s 28 -c 14 -o gpurun_out/prof_full python tools/prof_replay.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
