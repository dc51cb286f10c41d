#!/bin/bash
# A/B of every library build under _C/ on the fused policy rollout (with and without the critic head); run on the GPU box
cd "$(dirname "$0")/.."
for lib in autonomous_quadrotor_environment_b200/_C/libquadsim*.so; do
  echo "== $lib"
  QUADSIM_LIB=$PWD/$lib timeout 300 python tools/kbench.py policycritic
done
