import os, torch, pynvml
pynvml.nvmlInit()
uuid = "GPU-" + str(torch.cuda.get_device_properties(0).uuid)
print(uuid)
try:
    h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
except Exception as e:
    print("str failed", e); h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
allowed = os.sched_getaffinity(0)
print("allowed", len(allowed), sorted(allowed)[:4], "...", max(allowed), "cpu_count", os.cpu_count())
words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) + 64) // 64)
print("words", [hex(int(w)) for w in words])
try:
    print("numa", pynvml.nvmlDeviceGetNumaNodeId(h))
except Exception as e:
    print("numa id:", e)
import sys; sys.path.insert(0, os.getcwd())
from multiagent_gnn_policies_b200 import parallel
print("bind ->", parallel.bind_to_local_cpus(0))
print("now", len(os.sched_getaffinity(0)))
