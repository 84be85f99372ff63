set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_scan.py -x -q > gpurun_out/r2g_scan_tests.log 2>&1
timeout 600 python tools/probe_tri.py > gpurun_out/r2g_probe_tri.log 2>&1
timeout 600 python tools/probe_tri.py --timing > gpurun_out/r2g_probe_tri_timing.log 2>&1
timeout 600 python tools/probe_fp64.py > gpurun_out/r2g_fp64.log 2>&1
tail -4 gpurun_out/r2g_scan_tests.log; cat gpurun_out/r2g_probe_tri.log; grep -A1 "phase" gpurun_out/r2g_probe_tri_timing.log; tail -4 gpurun_out/r2g_fp64.log
