set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tools/probe_tri.py --timing > gpurun_out/r2i_probe_tri_timing.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_gpu_tests.log 2>&1
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
grep phase gpurun_out/r2i_probe_tri_timing.log; tail -3 gpurun_out/r2i_gpu_tests.log; tail -3 gpurun_out/r2i_bench.err; cat gpurun_out/r2i_bench.json | head -c 600
