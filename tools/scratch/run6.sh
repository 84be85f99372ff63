set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_scan.py -x -q > gpurun_out/r2f_scan_tests.log 2>&1
timeout 600 python tools/probe_tri.py > gpurun_out/r2f_probe_tri.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_gpu_tests.log 2>&1
tail -4 gpurun_out/r2f_scan_tests.log; cat gpurun_out/r2f_probe_tri.log; tail -4 gpurun_out/r2f_gpu_tests.log
