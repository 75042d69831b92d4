import torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    """GPU time per call: the n calls are captured in a CUDA graph so that host launch overhead is excluded."""
    side=torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3): f(i%n)
    torch.cuda.current_stream().wait_stream(side)
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n): f(i)
    g.replay()
    a,c=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for r in range(3): g.replay()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c)*1e3/(3*n)
print(os.environ.get('TAG',''), 'L1 fwd us %.1f' % timeit(lambda i: ops.cheb_fwd(xr[i],permt,*pl1.tensors(),W1,b,5,4,1,True,True,2),25),
      'L2 fwd us %.1f' % timeit(lambda i: ops.cheb_fwd(xs[i],None,*pl2.tensors(),W2,b,5,4,1,True,True,2),40))
