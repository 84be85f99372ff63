set -x
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2_strong.json 2> gpurun_out/r2e_bench_n2_strong.err
tail -5 gpurun_out/r2e_bench_n2_strong.err
cat gpurun_out/r2e_bench_n2_strong.json | head -c 3000
