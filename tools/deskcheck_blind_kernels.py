"""Desk checks of the kernels written without GPU access at the end of round 1 (gru_fwd_v2_kernel, conv1d_fwd_v2_kernel,
conv1d_dgrad_v2_kernel): numpy re-enactments of each kernel's THREAD MAPPING and index arithmetic (who owns which unit /
position, shared-memory layouts and paddings, packed-weight layouts, shuffle partners, tile bounds, truncating division)
against a direct evaluation of the operator.  They validate the index logic, not the CUDA: the GPU parity tests remain
the gate (tools/round2_sweep.py).  Run: python tools/deskcheck_blind_kernels.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import model_oracle as mo

rng = np.random.default_rng(0)


def gru_fwd_v2(H=64, L=7, I=16):
    """Thread (j, q) = (tid >> 1, tid & 1): unit j, K half q; h halves padded by 4 floats; (r, z) weights packed per k, n-gate
    weights packed over k pairs; one xor-1 shuffle; constants folded; stash slots by lane."""
    KH=H//2; HPAD=H+8
    rng=np.random.default_rng(0)
    x=rng.standard_normal((1,L,I)); w_ih=rng.standard_normal((3*H,I))*0.2; w_hh=rng.standard_normal((3*H,H))*0.2
    b_ih=rng.standard_normal(3*H)*0.1; b_hh=rng.standard_normal(3*H)*0.1
    out=mo.gru_direction(*[torch.from_numpy(a) for a in (x,w_ih,w_hh,b_ih,b_hh)], reverse=False, steps=None).numpy()[0]
    gi=(x[0]@w_ih.T+b_ih)   # [L,3H]
    C1=-1.4426950408889634; C2=2*C1
    hsm=np.zeros((2,HPAD)); hprev=np.zeros(2*H)
    # per-thread weights
    w_rz=np.zeros((2*H,KH,2)); w_n=np.zeros((2*H,KH//2,2)); b_rz=np.zeros((2*H,2)); b_n=np.zeros((2*H,2))
    for tid in range(2*H):
        j,q=tid>>1,tid&1
        for i in range(KH):
            w_rz[tid,i]=(C1*w_hh[0*H+j,q*KH+i], C1*w_hh[1*H+j,q*KH+i])
        for i in range(KH//2):
            w_n[tid,i]=(C2*w_hh[2*H+j,q*KH+2*i], C2*w_hh[2*H+j,q*KH+2*i+1])
        if q==0:
            b_rz[tid]=(C1*b_hh[j],C1*b_hh[H+j]); b_n[tid]=(C2*b_hh[2*H+j],0)
    hs=np.zeros((L,H)); stash=np.zeros((L,4*H))
    for s in range(L):
        cur=s&1
        sr=np.zeros(2*H); sz=np.zeros(2*H); sn=np.zeros(2*H)
        for tid in range(2*H):
            j,q=tid>>1,tid&1
            a0=b_rz[tid].copy(); a1=np.zeros(2); an=b_n[tid].copy()
            base=q*(KH+4)
            for i4 in range(KH//4):
                h4=hsm[cur,base+4*i4:base+4*i4+4]
                a0+=w_rz[tid,4*i4+0]*h4[0]; a1+=w_rz[tid,4*i4+1]*h4[1]
                an+=w_n[tid,2*i4+0]*np.array([h4[0],h4[1]])
                a0+=w_rz[tid,4*i4+2]*h4[2]; a1+=w_rz[tid,4*i4+3]*h4[3]
                an+=w_n[tid,2*i4+1]*np.array([h4[2],h4[3]])
            sr[tid]=a0[0]+a1[0]; sz[tid]=a0[1]+a1[1]; sn[tid]=an[0]+an[1]
        sr=sr+sr[np.arange(2*H)^1]; sz=sz+sz[np.arange(2*H)^1]; sn=sn+sn[np.arange(2*H)^1]
        newh=hsm[cur^1].copy()
        for tid in range(2*H):
            j,q=tid>>1,tid&1
            g_r=C1*gi[s,j]; g_z=C1*gi[s,H+j]; g_n=C2*gi[s,2*H+j]
            rg=1/(1+2**(sr[tid]+g_r)); zg=1/(1+2**(sz[tid]+g_z)); sg=1/(1+2**(rg*sn[tid]+g_n)); ng=2*sg-1
            hn=zg*(hprev[tid]-ng)+ng; hprev[tid]=hn
            newh[j+(j//KH)*4]=hn
            if q==0: hs[s,j]=hn; stash[s,j]=rg; stash[s,H+j]=zg
            else: stash[s,2*H+j]=ng; stash[s,3*H+j]=sn[tid]/C2
        hsm[cur^1]=newh
    err_h = np.abs(hs-out).max()
    # stash q check: W_hn h_prev + b_hn
    hp=np.vstack([np.zeros((1,H)),out[:-1]])
    qref=hp@w_hh[2*H:].T+b_hh[2*H:]
    err_q = np.abs(stash[:,3*H:]-qref).max()
    return err_h, err_q


def conv_fwd_v2(CO,KW,S,P,TL,NP,CI,Lin):
    Lout=(Lin+2*P-KW)//S+1
    x=rng.standard_normal((CI,Lin)); w=rng.standard_normal((CO,CI,KW))
    ref=np.zeros((CO,Lout))
    xp=np.pad(x,((0,0),(P,P)))
    for l in range(Lout):
        ref[:,l]=np.einsum('ock,ck->o',w,xp[:,S*l:S*l+KW])
    TPOS=TL*NP; SPAN=(TPOS-1)*S+KW
    y=np.full((CO,Lout),np.nan)
    for blk in range((Lout+TPOS-1)//TPOS):
        l0=blk*TPOS; in0=l0*S-P
        xs=np.zeros((CI,SPAN))
        for c in range(CI):
            for i in range(SPAN):
                gi=in0+i
                xs[c,i]=x[c,gi] if 0<=gi<Lin else 0.0
        ws=np.zeros((CI*KW,CO))
        for o in range(CO):
            for ck in range(CI*KW):
                ws[ck,o]=w[o,ck//KW,ck%KW]
        for tid in range(TL):
            acc=np.zeros((NP,CO))
            for c in range(CI):
                for k in range(KW):
                    for j in range(NP):
                        xv=xs[c,tid*S+j*(TL*S)+k]
                        acc[j]+=ws[c*KW+k]*xv
            for j in range(NP):
                l=l0+tid+j*TL
                if l<Lout: y[:,l]=acc[j]
    return np.nanmax(np.abs(y-ref)), np.isnan(y).sum()


def floor_div2(a): return a//2 if a>=0 else -((-a+1)//2)
def conv_dgrad_v2(CO,KW,S,P,TI,CPAD,NP,CI,Lin):
    Lout=(Lin+2*P-KW)//S+1
    dy=rng.standard_normal((CO,Lout)); w=rng.standard_normal((CO,CI,KW))
    # reference: dx[c,i] = sum_{o,k: S*l+k-P==i} w[o,c,k]*dy[o,l]
    ref=np.zeros((CI,Lin))
    for l in range(Lout):
        for k in range(KW):
            i=S*l+k-P
            if 0<=i<Lin: ref[:,i]+=w[:,:,k].T@dy[:,l]
    TPOS=TI*NP; NL=(TPOS-1+KW-1)//S+2
    dx=np.full((CI,Lin),np.nan)
    for blk in range((Lin+TPOS-1)//TPOS):
        i0=blk*TPOS; lbase=floor_div2(i0+P-(KW-1))
        dys=np.zeros((CO,NL))
        for o in range(CO):
            for ll in range(NL):
                l=lbase+ll
                dys[o,ll]=dy[o,l] if 0<=l<Lout else 0.0
        for tid in range(TI):
            ibase=i0+tid
            acc=np.zeros((NP,CPAD))
            k=(ibase+P)&1
            while k<KW:
                num=ibase+P-k
                ll0=int(num/S)-lbase   # C truncation
                for j in range(NP):
                    ll=ll0+j*(TI//S)
                    ok=0<=ll<NL
                    for o in range(CO):
                        d=dys[o,ll if ok else 0]*(1.0 if ok else 0.0)
                        acc[j,:CI]+=w[o,:,k]*d
                k+=S
            for j in range(NP):
                i=ibase+j*TI
                if i<Lin: dx[:,i]=acc[j,:CI]
    return np.nanmax(np.abs(dx-ref)), np.isnan(dx).sum()


if __name__ == "__main__":
    for H in (64, 32):
        eh, eq = gru_fwd_v2(H=H)
        print(f"gru_fwd_v2 H={H}: max |h - oracle| = {eh:.2e}, max |stashed q - reference| = {eq:.2e}")
        assert eh < 1e-12 and eq < 1e-12
    for args in ((16, 7, 2, 3, 128, 2, 6, 480), (32, 5, 2, 2, 64, 2, 16, 333), (16, 7, 2, 3, 128, 2, 3, 641)):
        err, holes = conv_fwd_v2(*args)
        print(f"conv1d_fwd_v2 {args}: max err {err:.2e}, unwritten outputs {holes}")
        assert err < 1e-12 and holes == 0
    for args in ((16, 7, 2, 3, 128, 8, 2, 6, 480), (32, 5, 2, 2, 64, 16, 2, 16, 333), (16, 7, 2, 3, 128, 8, 2, 3, 641)):
        err, holes = conv_dgrad_v2(*args)
        print(f"conv1d_dgrad_v2 {args}: max err {err:.2e}, unwritten outputs {holes}")
        assert err < 1e-12 and holes == 0
    print("desk checks ok")
