import sys, torch
sys.path.insert(0, '.')
from ball_action_spotting_b200 import _lib
import torch.nn.functional as F
lib = _lib.load()
DEV='cuda:0'
torch.manual_seed(0)
n,H,W,cin,cmid=2,368,640,32,16
x=(torch.randn(n,H,W,cin,device=DEV)*1.0).half()
w1=(torch.randn(cmid,9*cin,device=DEV)*(2/(9*cin))**0.5).half(); b1=torch.randn(cmid,device=DEV)*0.1
outs={}
for mode in (2,1,0):
    lib.mds_set_conv_mode(mode)
    out=torch.zeros(n,H,W,cmid,device=DEV).half()
    rc=lib.mds_k_conv3x3(x.data_ptr(),out.data_ptr(),w1.data_ptr(),b1.data_ptr(),None,None,n,H,W,cin,cmid,1,0,0,None)
    torch.cuda.synchronize(); assert rc==0
    outs[mode]=out.float()
wf=w1.float().view(cmid,3,3,cin).permute(0,3,1,2).contiguous()
ref=F.silu(F.conv2d(x.float().permute(0,3,1,2).double(), wf.double(), b1.double(), padding=1)).permute(0,2,3,1).float()
for mode in (2,1,0):
    d=(outs[mode]-ref)
    print('mode',mode,'rms err',d.pow(2).mean().sqrt().item(),'max',d.abs().max().item(),'mean',d.mean().item())
d=(outs[2]-outs[1])
nz=(d!=0)
print('fold vs nofold: differing',nz.sum().item(),'of',d.numel(),'max',d.abs().max().item())
xs=nz.any(dim=3).any(dim=1).any(dim=0).nonzero().flatten()
print('cols with diffs mod 30 histogram', torch.bincount(xs%30, minlength=30).tolist())
err2=(outs[2]-ref).abs(); err1=(outs[1]-ref).abs()
print('per col-mod-30 mean err fold', [round(v,7) for v in err2.mean(dim=(0,1,3)).view(-1)[:60].tolist()][:32])
print('per col mean err nofold    ', [round(v,7) for v in err1.mean(dim=(0,1,3)).view(-1)[:60].tolist()][:32])
