#!/bin/bash
# multi-GPU call: the sharded tests on real devices, then bench.py at N = $1 under torchrun (strong scaling)
N=${1:-2}; out=gpurun_out; mkdir -p $out
date -u +%T
nvidia-smi -L | head -8
[ -z "$SKIP_TESTS" ] && timeout 300 python -m pytest tests/test_gpu_shard.py "tests/test_gpu_dropin.py::test_cpp_sharded_host_path_equals_single_gpu" -m gpu -x -q > $out/r2m${N}_pytest.log 2>&1; tail -4 $out/r2m${N}_pytest.log
date -u +%T
for n in $2; do
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $out/r2m_bench_n1.json 2> $out/r2m_bench_n1.err
  else
    NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > $out/r2m_bench_n$n.json 2> $out/r2m_bench_n$n.err
  fi
  tail -2 $out/r2m_bench_n$n.err | cut -c1-300
  python - <<P
import json
try:
    j=json.load(open("gpurun_out/r2m_bench_n$n.json")); c=j["config"]
    print("N=%d value %.0f ms/step %.4f e2e %.0f allgather_ms %.4f default %.2f solve %.4f" % (j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["value"], c["allgather_ms"], j["e2e_default_criteria"]["ms_per_call_32_objects"], c["solve_ms"]))
    print("   traj", c["trajectory_ms_per_step"])
    print("   pass", c["pass_ms_per_step"])
except Exception as e: print("no line", e)
P
done
date -u +%T
