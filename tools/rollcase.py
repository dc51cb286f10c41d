import sys, os
sys.path.insert(0, "/root/repo/tools")
import kbench
for S in (1, 2, 4, 8):
    kbench.case_rollout(n=10 ** 9, substeps=S, K=32, iters=6)
    kbench.case_rollout(async_reset=True, T=5, substeps=S, K=32, iters=6)
