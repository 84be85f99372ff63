set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-sample > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_lb -s 1 -c 1 -f -o gpurun_out/r2_lb_final python tools/ncu_scan_once.py > gpurun_out/r2_ncu_lb_final.log 2>&1
ncu --metrics sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_fp64_ncu.csv python tools/probe_fp64.py --once > gpurun_out/r2_fp64_once.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_scan.py > gpurun_out/r2_san_memcheck.log 2>&1
tail -3 gpurun_out/r2_san_memcheck.log; tail -2 gpurun_out/r2_ncu_lb_final.log; wc -l gpurun_out/r2_launches.csv gpurun_out/r2_fp64_ncu.csv
