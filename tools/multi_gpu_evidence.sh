#!/bin/bash
# usage (through gpurun --gpus N): tools/multi_gpu_evidence.sh N   — host-link ceiling with N ranks copying at once, the bench lines
# (driver arguments and a long run) and the two-device test
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 tools/pcie_bw_ranks.py > gpurun_out/r02_pcie_bw_${N}ranks.txt 2>&1
tail -n 40 gpurun_out/r02_pcie_bw_${N}ranks.txt
timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}_driverargs.json 2> gpurun_out/r02_bench_n${N}_driverargs.err
timeout 300 $TR --master-port 29523 bench.py --gpus $N --steps 2000 --warmup 50 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
python - <<PY
import json
for f in ("r02_bench_n${N}_driverargs", "r02_bench_n${N}"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, d["value"], d["n_gpus"], d["ms_per_step"], d["e2e"]["value"], d["stats"].get("exchanges_in_timed_region"), d["clocks"])
PY
timeout 200 python -m pytest tests -m gpu -q -k "second_device" 2>&1 | tail -n 2
