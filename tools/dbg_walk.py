import sys, numpy as np
sys.path.insert(0, ".")
from changa_b200.hostcuda import HostCUDA
from changa_b200.workloads import config_workload
from changa_b200.device_step import DeviceTreeStep
hc = HostCUDA(double=False, device=0)
wl = config_workload("cube300", n=14 ** 3)
t = wl["tree"]
step = DeviceTreeStep(hc, t, theta=0.7, n_replicas=1, period=1.0, ewald=None)
step.run(keep_lists=True)
print(step.lists_info)
