"""TEST INFRASTRUCTURE (GPU box).  Per-layer gradient sums and open ReLU-gate counts of one sparse-encoder training step, the reference module
(oracle/ref_shim.py + oracle/spconv_shim.py, CPU) next to the CUDA encoder (SHERF_SP_DEBUG=1 prints its lines on stderr): a unit whose
pre-activation is zero to rounding can open on one side only, which shows here as a gate count that differs by one and explains a ~ 1 / rows
jump of every gradient below that layer.  usage: python tests/helpers/sparse_gate_counts.py [train]"""
import os, sys, torch
os.environ['SHERF_SP_DEBUG']='1'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch.nn.functional as F
from oracle import ref_shim, sparse_encoder as SE
from sherf_b200 import synthetic as S
from sherf_b200.renderer import SparseConvNet, SparseConvTensor
from test_sparse_encoder import _shell
m = S.make_smpl_model(0)
rr,_ = ref_shim.load(S.smpl_model_to_torch(m))
import spconv
dev=torch.device('cuda:0')
shape=(32,64,64); n=200; dup=10
torch.manual_seed(0)
ref = rr.SparseConvNet(num_layers=4); sd = SE.random_state_dict(ref, 7); ref.load_state_dict(sd)
ours = SparseConvNet(4); ours.load_state_dict(sd)
coord, feat = _shell(n, shape, 13, dup=dup)
idx = torch.cat([torch.zeros(coord.shape[0],1,dtype=torch.int32), coord],1)
g = torch.Generator().manual_seed(17)
grid = (torch.rand(1,1,1,900,3,generator=g)*2-1)*0.95
cot = torch.randn(1,900,192,generator=g)
train = len(sys.argv) > 1 and sys.argv[1]=='train'
ref.train(train).requires_grad_(True)
relus=[]
for name in ['conv0','down0','conv1','down1','conv2','down2','conv3']:
    blk=getattr(ref,name)
    for i in range(2,len(blk),3): relus.append((name+'.'+str(i), blk[i]))
caps={}
for nm,mod in relus:
    def hk(mod, gin, gout, nm=nm):
        caps[nm]=(gout[0].detach().clone())
    mod.register_full_backward_hook(hk)
    def fh(mod, inp, out, nm=nm):
        caps['act/'+nm]=out.detach().clone()
    mod.register_forward_hook(fh)
f = feat.clone().requires_grad_(True)
out = ref(spconv.core.SparseConvTensor(f, idx, list(shape), 1), grid)
(out*cot).sum().backward()
for c,(nm,_) in enumerate(relus):
    dA=caps[nm].double(); act=caps['act/'+nm]
    print(f'[ref bwd]  layer {c:2d} rows {dA.shape[0]:4d} sum dA {float(dA.sum()): .6e}  sum|dA| {float(dA.abs().sum()):.6e}  gated sum {float(dA[act>0].sum()): .6e}  open gates {int((act>0).sum())}')
ours = ours.to(dev).train(train).requires_grad_(True)
fo = feat.to(dev).requires_grad_(True)
vols = ours(SparseConvTensor(fo, idx.to(dev), list(shape), 1))
feats = torch.cat([F.grid_sample(v, grid.to(dev), padding_mode='zeros', align_corners=True) for v in vols], dim=1)
oo = feats.view(1,-1,feats.size(4)).transpose(1,2)
(oo*cot.to(dev)).sum().backward()
torch.cuda.synchronize()
