#!/bin/bash
# A/B of every library build under _C/ (DESIGN.md "Tuning builds") on the headline step case and the fused rollouts; run on the GPU box
cd "$(dirname "$0")/.."
for lib in autonomous_quadrotor_environment_b200/_C/libquadsim*.so; do
  echo "== $lib"
  QUADSIM_LIB=$PWD/$lib timeout 200 python tools/kcase.py sensor_noise=1 async_reset=1 T=5 iters=2000
  QUADSIM_LIB=$PWD/$lib timeout 200 python tools/kcase.py sensor_noise=0 async_reset=1 T=5 iters=2000
  [ "$1" = "norollout" ] || QUADSIM_LIB=$PWD/$lib timeout 300 python tools/kbench.py rolloutab
done
