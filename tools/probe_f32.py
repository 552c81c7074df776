import sys; sys.path.insert(0, "nbody-gnn-hpc_b200")
import torch, numpy as np
from hpc import _cuda, ics
eng=_cuda.get_engine()
for n in (65536, 262144):
    x,_,m=ics.plummer_ic(n,seed=7)
    pos_d=eng.to_device(x); m_d,f32=eng._masses_dev(m)
    stream=eng.pack(pos_d,m_d,f32,n,np.float32); ws=eng.workspace(n,n,np.float32)
    for _ in range(3): eng.accel_slab(stream,n,0,n,0.01,ws)
    best=1e9
    for rep in range(3):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10 if n<100000 else 3): eng.accel_slab(stream,n,0,n,0.01,ws)
        e1.record(); e1.synchronize()
        best=min(best,e0.elapsed_time(e1)/(10 if n<100000 else 3))
    gi=n*(n-1.0)/best/1e6
    print(n, round(best,4),"ms", round(gi,1),"G/s", round(gi*20/74450,4))
