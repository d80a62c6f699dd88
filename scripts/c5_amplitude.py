import sys, time
sys.path.insert(0,'.')
import torch, numpy as np
import tedq_b200 as qb
from tedq_b200 import workloads as W
rows,cols,cyc=5,8,12
spec=W.lattice_rcs(rows,cols,cyc,seed=0)
circ=W.build_circuit(spec,qb)
t=time.time()
cc=circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt={"max_repeats":int(sys.argv[1]) if len(sys.argv)>1 else 64,"slicing_opts":{"target_size":2**27,"target_num_slices":64}})
print('compile',time.time()-t)
bits=[0]*40
amp=cc.amplitude(bits); torch.cuda.synchronize()
plan=cc._tn._amplitude_plan()[2]; info=cc._tn._amplitude_plan()[1]
print(info, 'flops/slice %.3e'%plan.flops, 'slices',plan.n_slices,'width',plan.width,'steps',plan.n_steps)
for _ in range(2):
    torch.cuda.synchronize(); t=time.time()
    amp=cc.amplitude(bits); torch.cuda.synchronize(); dt=time.time()-t
    print('amp',complex(amp.cpu()),'time',dt,'TFLOP/s',plan.flops*plan.n_slices/dt/1e12)
# slice invariance
a=cc.amplitude(bits, slice_range=(0,32))+cc.amplitude(bits, slice_range=(32,64))
print('halves', complex(a.cpu()))
# histogram of step sizes
import collections
h=collections.Counter()
big=[]
for s in range(plan.n_steps):
    st=plan.step(s); k,m,n,b=st[2],st[3],st[4],st[5]
    fl=8*2.0**(k+m+n+b)
    if fl>plan.flops*0.01: big.append((fl/plan.flops,k,m,n,b))
print('dominant steps (share,k,m,n,b):',sorted(big,reverse=True)[:12])
