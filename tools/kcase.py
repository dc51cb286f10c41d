"""Developer tool: one kbench step case from the command line, e.g.
    python tools/kcase.py sensor_noise=1 async_reset=1 T=5 act_scale=1.0 n=1000 iters=50"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kbench

kw = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    kw[k] = (v if k in ("precision", "integrator") else (float(v) if "." in v else int(v)))
for k in ("sensor_noise", "async_reset", "auto_reset"):
    if k in kw:
        kw[k] = bool(kw[k])
kbench.case_step(**kw)
