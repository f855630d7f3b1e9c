"""NumPy emulation of the pair kernel's nearest-neighbour reuse thresholds on the configs[2] batch (exact budgets):
how many 32-source slot decisions a perfect bound would need per iteration.  Design study, CPU only."""
import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import icp_oracle as orc
P=24
src,tgt=orc.synth_room_batch(0,P)
src=src.astype(np.float64); tgt=tgt.astype(np.float64)
def fit(A,B):
    ca=A.mean(0); cb=B.mean(0); H=(A-ca).T@(B-cb)
    num=H[0,1]-H[1,0]; den=H[0,0]+H[1,1]; h=np.hypot(num,den); c,s=den/h,num/h
    R=np.array([[c,-s],[s,c]]); t=cb-R@ca; return R,t
tot_dec=np.zeros(30); tot_dec_perslot=np.zeros(30); tot_dec_persrc=np.zeros(30)
mvs=np.zeros((P,30)); buds=[]
for p in range(P):
    S=src[p].copy(); T=tgt[p]; c=T.mean(0)
    smax=np.abs(S-c).max()*1.4142137
    cum=0.0
    npass=6; tnn=np.full(npass,-np.inf); tnn_slot=np.full(12,-np.inf); tnn_src=np.full(360,-np.inf)
    for it in range(30):
        D=np.sqrt(((S[:,None,:]-T[None,:,:])**2).sum(2))
        o=np.argsort(D,axis=1); d1=D[np.arange(360),o[:,0]]; d2=D[np.arange(360),o[:,1]]
        bud=0.5*(d2-d1)
        for q in range(npass):
            if not cum<=tnn[q]:
                tot_dec[it]+= min(64,360-64*q)/32
                b=bud[64*q:64*q+64].min(); tnn[q]=cum+0.999*b
                if it==8: buds.append(b)
        for q in range(12):
            if not cum<=tnn_slot[q]:
                tot_dec_perslot[it]+=1
                tnn_slot[q]=cum+0.999*bud[32*q:32*q+32].min()
        # per-source idealised
        need=~(cum<=tnn_src)
        tot_dec_persrc[it]+=need.sum()/32
        tnn_src[need]=cum+0.999*bud[need]
        R,t=fit(S,T[o[:,0]])
        S2=S@R.T+t
        rho=np.hypot(R[0,0]-1,R[1,0]); dd=np.linalg.norm((R-np.eye(2))@c+t)
        mv=rho*smax+dd
        mvs[p,it]=mv; actual=np.linalg.norm(S2-S,axis=1).max()
        smax+=mv; cum+=mv; S=S2
print("slot-decisions per pair-iteration (per-pass thresholds):", (tot_dec/P).round(2), "mean", tot_dec.sum()/P/30)
print("per-slot thresholds:", (tot_dec_perslot/P).round(2), tot_dec_perslot.sum()/P/30)
print("per-source ideal:", (tot_dec_persrc/P).round(2), tot_dec_persrc.sum()/P/30)
print("median mv per iteration:", np.median(mvs,axis=0))
print("budgets at it 8: median", np.median(buds), "10%", np.percentile(buds,10))
