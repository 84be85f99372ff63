set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2j_n2_fused.json 2> gpurun_out/r2j_n2_fused.err
timeout 600 env DIG_NO_MULTICAST=1 $TR --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2j_n2_unicast.json 2> gpurun_out/r2j_n2_unicast.err
timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --exchange nccl > gpurun_out/r2j_n2_nccl.json 2> gpurun_out/r2j_n2_nccl.err
for f in fused unicast nccl; do tail -4 gpurun_out/r2j_n2_$f.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2j_n2_$f.json"))
    print("$f", d["ms_per_step"], d["roofline"]["kernel_ms"], d["parity_sample"], d["sharding"]["exchange"][:160])
except Exception as e: print("$f failed", e)
PY
done
