"""Developer tool: fused rollout (qs_rollout) at several RK4 sub-interval counts, with and without asynchronous resets."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kbench
for S in [int(x) for x in (sys.argv[1:] or ["1", "2", "4", "8"])]:
    kbench.case_rollout(n=10 ** 9, substeps=S, K=32, iters=6)
    kbench.case_rollout(async_reset=True, T=5, substeps=S, K=32, iters=6)
