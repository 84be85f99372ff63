set -x
cd $GRAFT_REPO_ROOT
python tools/probe_lb.py --timing > gpurun_out/r2c_probe_timing.log 2>&1
python tools/scratch/dbg_det2.py > gpurun_out/r2c_det2.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_gpu_tests.log 2>&1
python bench.py --steps 10 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -5 gpurun_out/r2c_gpu_tests.log; tail -6 gpurun_out/r2c_det2.log; tail -5 gpurun_out/r2c_bench.err; grep -E "fused53|process|wait|write-out|barrier" gpurun_out/r2c_probe_timing.log
