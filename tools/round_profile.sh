#!/bin/bash
# Round evidence on one B200 (run through gpurun): the bench line, the ncu launch list of the same command, one `--set full`
# capture of the headline kernel and of the fused sensor rollout, DRAM traffic for bench.py's roofline.traffic.
cd "$(dirname "$0")/.."
R=${1:-r02}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 40 --warmup 3 --variant-steps 0 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${R}_launches_bench.log 2>&1
timeout 300 tools/prof2.sh ${R}_prof_step_pair_sensor "step_kernel_pair" 30 16384 profsensor > /dev/null 2>&1
timeout 300 tools/prof2.sh ${R}_prof_policy_critic "policy_rollout_kernel" 1 2048 profpolicycritic > /dev/null 2>&1
timeout 200 python tools/kbench.py policycritic > gpurun_out/${R}_kbench_policy.txt 2>&1
timeout 300 tools/prof2.sh ${R}_prof_ppo_grad "ppo_grad_kernel" 1 16384 profppograd > /dev/null 2>&1
timeout 200 python tools/kbench.py ppograd > gpurun_out/${R}_kbench_ppo_grad.txt 2>&1
timeout 200 python tools/kbench.py f64 > gpurun_out/${R}_kbench_f64.txt 2>&1
timeout 300 python tools/train_ppo.py --envs 65536 --iters 4 > gpurun_out/${R}_train_ppo.txt 2>&1
ls -la gpurun_out | tail -n 20
