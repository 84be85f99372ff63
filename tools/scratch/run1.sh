set -x
cd $GRAFT_REPO_ROOT
python tools/probe_lb.py --timing > gpurun_out/r2b_probe_timing.log 2>&1
python tools/probe_lb.py --repeat 24 --cold > gpurun_out/r2b_probe_repeat.log 2>&1
python tools/scratch/dbg_det2.py > gpurun_out/r2b_det2.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.log 2>&1
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_gpu_tests.log; tail -5 gpurun_out/r2b_probe_repeat.log; tail -4 gpurun_out/r2b_det2.log; cat gpurun_out/r2b_bench.err | tail -5
