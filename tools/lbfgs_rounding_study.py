"""Where does the device L-BFGS's distance from float64 come from?  (test infrastructure: executes oracle/ on the CPU)

Records the gradients of a 160-evaluation reference run (64^2, fp32 oracle), then replays them ("teacher forcing") through
  * oracle.LbfgsRestated in float64 (the yardstick) and in float32 tensors (what torch.optim.LBFGS itself computes), and
  * an emulation of csrc/lbfgs_impl.cuh's rearranged recursion (dot-product matrices SY / YY + O(m^2) scalar solve + one
    accumulation pass for d) with (a) exact float64 dots of the fp32-stored history or (b) fp32 sums inside 512-element tiles
    as the kernel forms them, and with the direction accumulated in fp32 (as shipped), fp64, reversed order or Kahan-compensated.
Every variant keeps its OWN history (s = t * d uses its own d), exactly like the real optimisers, so direction errors feed back.
Result (profiles/r02_lbfgs_rounding_study.log): torch-style fp32 is 9e-6 from float64 at worst; the shipped arithmetic 3-6e-5;
exact dots + fp64 accumulation 9e-6 — i.e. ~1e-5 is the floor set by storing the history in fp32 as torch does.
"""
import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ist_oracle as O, synth
torch.set_num_threads(8)
size=64
state_np = synth.vgg_state_dict(0, upto="conv5_1")
st = O.state_to_torch(state_np, torch.float32)
content = torch.from_numpy(synth.preprocess(synth.smooth_frame(size,1)))
style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size,2)))
targets = O.compute_targets(st, content, style, full=False)
x = content.clone().requires_grad_(True)
opt = torch.optim.LBFGS([x])
rec=[]
n=[0]
while n[0] < 160:
    def closure():
        opt.zero_grad()
        loss = sum(O.layer_losses(st, x, targets, full=False)); loss.backward(); n[0]+=1
        rec.append((float(loss), x.grad.detach().flatten().clone()))
        return loss
    opt.step(closure)
print('recorded', len(rec))
# teacher-forced fp64 reference
def replay(dtype):
    r = O.LbfgsRestated(); r.log=[]
    xr = content.flatten().to(dtype).clone()
    k=[0]
    for step in range(8):
        def closure():
            l,g = rec[k[0]]; k[0]+=1
            return l, g.to(dtype).clone()
        r.step(xr, closure)
    return r.log
log64 = replay(torch.float64); log32 = replay(torch.float32)
print('fp32 torch-style worst', max(float((a['d'].double()-b['d']).norm()/b['d'].norm()) for a,b in zip(log32,log64)))
# expanded algorithm (device) emulation
def expanded(dot):
    m=100
    S=[];Y=[];ro=[]
    H=1.0; d=None; t=None; prev_g=None
    out=[]
    for it,(l,g) in enumerate(rec):
        g = g.float()
        if it==0:
            d = -g.double(); 
        else:
            y = (g - prev_g)            # fp32
            s = (torch.tensor(t,dtype=torch.float32)*d32)   # fp32
            ys = dot(y,s)
            if ys>1e-10:
                if len(S)==m: S.pop(0);Y.pop(0);ro.pop(0)
                S.append(s);Y.append(y);ro.append(1.0/ys); H = ys/dot(y,y)
            h=len(S)
            SY=np.array([[dot(S[i],Y[j]) for j in range(h)] for i in range(h)]) if False else None
            # build via matrix products in the chosen dot arithmetic (vectorised)
            if h==0:
                d=(torch.tensor(-H,dtype=torch.float32)*g).double(); d32=d.float(); out.append(d.clone()); prev_g=g.clone(); t=1.0; continue
            Sm=torch.stack(S); Ym=torch.stack(Y)
            SY=dot.mat(Sm,Ym); YY=dot.mat(Ym,Ym); sg=dot.mv(Sm,g); yg=dot.mv(Ym,g)
            al=np.zeros(h); 
            for i in range(h-1,-1,-1):
                u = sum(al[j]*SY[i,j] for j in range(i+1,h))
                al[i]=ro[i]*(-sg[i]-u)
            v = H*(-yg - YY@al)
            c=np.zeros(h); w=np.zeros(h)
            for i in range(h):
                wi = sum(c[j]*SY[j,i] for j in range(i))
                c[i]=al[i]-ro[i]*(v[i]+wi)
            # d = -H g + sum(-H al_j y_j + c_j s_j) in fp32 fma like the device
            if MODE=='f32':
                acc = (torch.tensor(-H,dtype=torch.float32)*g)
                for j in range(h):
                    acc = acc + torch.tensor(-H*al[j],dtype=torch.float32)*Y[j] + torch.tensor(c[j],dtype=torch.float32)*S[j]
                d = acc.double()
            elif MODE=='f32_c2':     # fp32 accumulation, every coefficient as an fp32 hi + lo pair
                def two(v):
                    hi=torch.tensor(v,dtype=torch.float32); lo=torch.tensor(v-float(hi),dtype=torch.float32); return hi,lo
                h0,l0=two(-H); acc = h0*g; acc = acc + l0*g
                for j in range(h):
                    a1,a2=two(-H*al[j]); b1,b2=two(c[j])
                    acc = acc + a1*Y[j]; acc = acc + a2*Y[j]; acc = acc + b1*S[j]; acc = acc + b2*S[j]
                d = acc.double()
            elif MODE=='f64':
                acc = -H*g.double()
                for j in range(h):
                    acc = acc + (-H*al[j])*Y[j].double() + c[j]*S[j].double()
                d = acc
            elif MODE=='f32coef64':   # fp32 accumulate, coefficients rounded to fp32 (as now) but reversed order
                acc = (torch.tensor(-H,dtype=torch.float32)*g)
                for j in range(h-1,-1,-1):
                    acc = acc + torch.tensor(-H*al[j],dtype=torch.float32)*Y[j] + torch.tensor(c[j],dtype=torch.float32)*S[j]
                d = acc.double()
            elif MODE=='kahan':
                acc = (torch.tensor(-H,dtype=torch.float32)*g); comp=torch.zeros_like(acc)
                for j in range(h):
                    for term in (torch.tensor(-H*al[j],dtype=torch.float32)*Y[j], torch.tensor(c[j],dtype=torch.float32)*S[j]):
                        yk = term - comp; tk = acc + yk; comp = (tk - acc) - yk; acc = tk
                d = acc.double()
        d32 = d.float()
        out.append(d.clone())
        prev_g = g.clone()
        t = min(1.0, 1.0/float(g.double().abs().sum())) if it==0 else 1.0
    return out
class D64:
    def __call__(self,a,b): return float(a.double().dot(b.double()))
    def mat(self,A,B): return (A.double()@B.double().T).numpy()
    def mv(self,A,g): return (A.double()@g.double()).numpy()
class DTile:
    # fp32 products+sums inside 512-element tiles (torch fp32 sum per tile), fp64 across tiles
    def _t(self,P):  # P: [..., n] fp32 products
        n=P.shape[-1]; pad=(-n)%512
        if pad: P=torch.nn.functional.pad(P,(0,pad))
        return P.view(*P.shape[:-1],-1,512).sum(-1,dtype=torch.float32).double().sum(-1)
    def __call__(self,a,b): return float(self._t(a*b))
    def mat(self,A,B): return torch.stack([self._t(A[i][None,:]*B) for i in range(A.shape[0])]).numpy()
    def mv(self,A,g): return self._t(A*g[None,:]).numpy()
import itertools
class DLane16(DTile):
    # fp32 sums over the 16 elements one lane holds, float64 from there on (shuffle tree and across tiles in double)
    def _t(self,P):
        n=P.shape[-1]; pad=(-n)%16
        if pad: P=torch.nn.functional.pad(P,(0,pad))
        return P.view(*P.shape[:-1],-1,16).sum(-1,dtype=torch.float32).double().sum(-1)
for (name,dd),MODE in itertools.product((('exact fp64 dots of fp32 history',D64()),('fp32 in-tile dots',DTile()),('fp32 per-lane (16 elements) dots',DLane16())),('f32','f64','f32_c2')):
    out = expanded(dd); name=name+' / d accumulation '+MODE
    errs=[float((a-b['d']).norm()/b['d'].norm()) for a,b in zip(out,log64)]
    print(name,'worst',max(errs),'at',int(np.argmax(errs)), 'median', float(np.median(errs)))
