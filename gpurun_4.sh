timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import torch, os, sys, itertools
sys.path.insert(0,'.')
from gcn_fmri_decoding_b200 import ops, synth
from gcn_fmri_decoding_b200.plan import GraphPlan
dev=torch.device('cuda:0')
A,gs,perm,L=synth.brain_graph(4)
pl1=GraphPlan(L[0],dev); pl2=GraphPlan(L[2],dev)
permt=torch.as_tensor(perm,dtype=torch.int32,device=dev)
xr=[torch.randn(512,360,15,device=dev) for _ in range(25)]
xs=[torch.randn(512,100,32,device=dev) for _ in range(40)]
W1=torch.randn(75,32,device=dev)*.2; W2=torch.randn(160,32,device=dev)*.2; b=torch.full((32,),.2,device=dev)
def timeit(f,n):
    for i in range(5): f(i%n)
    a,c=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for r in range(3):
        for i in range(n): f(i)
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c)*1e3/(3*n)
print('L1 fwd us', timeit(lambda i: ops.cheb_fwd(xr[i],permt,*pl1.tensors(),W1,b,5,4,1,True,True,2),25))
print('L2 fwd us', timeit(lambda i: ops.cheb_fwd(xs[i],None,*pl2.tensors(),W2,b,5,4,1,True,True,2),40))
PY
