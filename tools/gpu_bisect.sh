set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PROF_REPS=2
for k in k_ingest_bulk k_normals k_sort_hist k_sort_scatter k_gather_points k_segment_heads k_score_work k_score_coop; do
  timeout 600 ncu --set full --clock-control none -k regex:"$k" -s 1 -c 8 -o gpurun_out/bisect_$k -f python tools/prof_replay.py 2>&1 | grep -E "^variant" | sed "s/^/[$k] /"
done
rm -f gpurun_out/bisect_*.ncu-rep
