#!/bin/bash
# usage (on the GPU box): tools/sanitize.sh [cases...]   — compute-sanitizer memcheck + synccheck over tools/sanitize_case.py.
# PYTORCH_NO_CUDA_MEMORY_CACHING=1: every torch tensor is its own cudaMalloc, so an out-of-bounds access is not hidden inside
# the caching allocator's segments.  Summaries land in gpurun_out/sanitize_<tool>.log.
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
for tool in memcheck synccheck; do
  timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $tool \
      --print-limit 20 python tools/sanitize_case.py "$@" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ok|ERROR SUMMARY|Invalid|Error|hazard|Barrier" gpurun_out/sanitize_$tool.log | tail -40
done
