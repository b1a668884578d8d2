import os, time, sys
sys.path.insert(0, os.getcwd())
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
print(open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0])
import torch
from oracle import cyclegan_oracle as O
t=time.time(); P = O.init_params(0, torch.float32); print("init", time.time()-t, flush=True)
A,B = O.synthetic_batch(0, 2, 128)
for th in (8, 16, 32, 64, os.cpu_count()):
    torch.set_num_threads(th)
    t=time.time()
    with torch.no_grad(): y = O.generator_forward(A, P, "generator_A2B")
    t1=time.time()-t
    t=time.time(); L,G,_,_ = O.gradients(A,B,P,10.0,5.0); t2=time.time()-t
    print("threads", th, "gen fwd B=2: %.3f s  full fwd+bwd B=2: %.2f s" % (t1, t2), flush=True)
