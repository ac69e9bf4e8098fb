import sys, numpy as np
sys.path.insert(0, "/root/repo")
from swiftest_b200 import Context, workloads as W
from swiftest_b200.context import FAM_PLPL
n = 100000
d = W.disk(n, seed=3031179)
mass = d["Gmass"] / W.GMSUN
with Context(0) as c:
    c.enable_kernel_timing(True)
    ms = []
    for it in range(6):
        pe = c.util_get_potential_energy(n, None, W.GMSUN, d["Gmass"], mass, d["rh"])
        ms.append(c.last_kernel_ms(FAM_PLPL))
    print("pe", pe, "kernel ms", np.round(ms, 3))
