#!/bin/bash
# A/B of the sensor step / rollout kernels across tuning builds (DESIGN.md "Tuning builds"); run on the GPU box
cd "$(dirname "$0")/.."
for lib in autonomous_quadrotor_environment_b200/_C/libquadsim*.so; do
  echo "== $lib"
  QUADSIM_LIB=$PWD/$lib timeout 200 python tools/kcase.py sensor_noise=1 async_reset=1 T=5 iters=2000
  QUADSIM_LIB=$PWD/$lib timeout 200 python tools/kbench.py ${1:-sensorrollout} | head -n 3
done
