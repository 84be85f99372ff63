set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_scan.py -x -q > gpurun_out/r2l_scan_tests.log 2>&1
timeout 600 python tools/scratch/dbg_det2.py > gpurun_out/r2l_det2.log 2>&1
timeout 600 python tools/probe_lb.py > gpurun_out/r2l_probe.log 2>&1
timeout 600 python tools/probe_lb.py --timing > gpurun_out/r2l_probe_timing.log 2>&1
timeout 600 python tools/probe_tri.py > gpurun_out/r2l_probe_tri.log 2>&1
timeout 600 python tools/probe_tri.py --timing > gpurun_out/r2l_probe_tri_timing.log 2>&1
tail -3 gpurun_out/r2l_scan_tests.log; grep "bad runs" gpurun_out/r2l_det2.log; grep -E "lane-bank|equal" gpurun_out/r2l_probe.log; sed -n 4,16p gpurun_out/r2l_probe_timing.log; cat gpurun_out/r2l_probe_tri.log; grep "phase" gpurun_out/r2l_probe_tri_timing.log
