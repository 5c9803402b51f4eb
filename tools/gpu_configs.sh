# usage: bash tools/gpu_configs.sh <configs...>   (on the GPU box; every step under its own timeout)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_cfg.log 2>&1; tail -6 gpurun_out/pytest_cfg.log
for c in "$@"; do
  timeout 600 python tools/run_configs.py $c > gpurun_out/cfg_$c.log 2>&1; tail -c 1500 gpurun_out/cfg_$c.log
done
