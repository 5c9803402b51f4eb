set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_staging_gpu.py tests/test_parity_gpu.py tests/test_host_cpp.py -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -8 gpurun_out/pytest_$tag.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err
cat gpurun_out/bench_n2_$tag.json; tail -20 gpurun_out/bench_n2_$tag.err
