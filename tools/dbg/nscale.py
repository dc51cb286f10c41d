import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.argv = [sys.argv[0]]
import importlib.util
spec = importlib.util.spec_from_file_location("kbench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kbench.py"))
kb = importlib.util.module_from_spec(spec)
sys.argv = [sys.argv[0], "none"]
spec.loader.exec_module(kb)
for N in (148 * 8 * 64 * 2, 148 * 8 * 64 * 4, 148 * 8 * 64 * 8, 148 * 8 * 64 * 14):
    kb.case_step(N=N, async_reset=True, T=5, sensor_noise=True, iters=400)
    kb.case_step(N=N, async_reset=True, T=5, sensor_noise=False, iters=400)
