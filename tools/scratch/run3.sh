set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_scan.py -x -q > gpurun_out/r2d_scan_tests.log 2>&1
python tools/scratch/dbg_det2.py > gpurun_out/r2d_det2.log 2>&1
python tools/probe_lb.py > gpurun_out/r2d_probe.log 2>&1
DIG_LIB_PATH=$GRAFT_REPO_ROOT/digdriver_b200/libdigb200_perlane.so python -m pytest tests/test_gpu_scan.py -x -q > gpurun_out/r2d_scan_tests_perlane.log 2>&1
DIG_LIB_PATH=$GRAFT_REPO_ROOT/digdriver_b200/libdigb200_perlane.so python tools/scratch/dbg_det2.py > gpurun_out/r2d_det2_perlane.log 2>&1
DIG_LIB_PATH=$GRAFT_REPO_ROOT/digdriver_b200/libdigb200_perlane.so python tools/probe_lb.py > gpurun_out/r2d_probe_perlane.log 2>&1
python tools/probe_lb.py --timing > gpurun_out/r2d_probe_timing.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_gpu_tests.log 2>&1
python bench.py --steps 10 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python tools/probe_fp64.py > gpurun_out/r2d_fp64.log 2>&1
tail -3 gpurun_out/r2d_scan_tests.log; grep "bad runs" gpurun_out/r2d_det2.log; grep -E "lane-bank|equal" gpurun_out/r2d_probe.log
tail -3 gpurun_out/r2d_scan_tests_perlane.log; grep "bad runs" gpurun_out/r2d_det2_perlane.log; grep -E "lane-bank|equal" gpurun_out/r2d_probe_perlane.log
tail -5 gpurun_out/r2d_gpu_tests.log; tail -5 gpurun_out/r2d_bench.err; cat gpurun_out/r2d_fp64.log
