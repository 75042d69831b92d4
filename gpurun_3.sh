timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:k_cheb_fwd_fused -s 2 -c 2 -o gpurun_out/prof_fwd2 python tools/prof_layers.py fwd 2 > gpurun_out/prof_fwd2.log 2>&1; tail -3 gpurun_out/prof_fwd2.log
for cfg in "1 4" "2 2" "2 1" "4 1"; do set -- $cfg; echo "L2 WS=$1 SG=$2"; GCNB_FWD_WS=$1 GCNB_FWD_SG=$2 python - <<'PY'
import torch, os, sys
sys.path.insert(0,'.')
from gcn_fmri_decoding_b200 import ops, synth
from gcn_fmri_decoding_b200.plan import GraphPlan
dev=torch.device('cuda:0')
A,gs,perm,L=synth.brain_graph(4)
pl2=GraphPlan(L[2],dev)
xs=[torch.randn(512,100,32,device=dev) for _ in range(40)]
W=torch.randn(160,32,device=dev)*.2; b=torch.full((32,),.2,device=dev)
f=lambda i: ops.cheb_fwd(xs[i],None,*pl2.tensors(),W,b,5,4,1,True,True,2)
for i in range(5): f(i)
a,c=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for r in range(3):
    for i in range(40): f(i)
c.record(); torch.cuda.synchronize()
print('us per call', a.elapsed_time(c)*1e3/120)
PY
done
