#!/usr/bin/env python
"""CPU model of lr_mark_marginal() (freesasa_b200/csrc/integrate.cu): the slices in which some pair of circles is within
rounding distance of a tangency are found in closed form — the slice plane touches the intersection circle of the two
SPHERES at z± = t dz/|D| ± rho dxy/|D| — instead of testing q = min(|N|,|D|)/max(|N|,|D|) < q_min for every pair of every
slice.  This script checks the construction against that brute-force test on the surface atoms of the 100k-atom globule
(PDB-rounded coordinates): the closed form must mark a SUPERSET.  numpy + the oracle's neighbour list; no GPU.

    python tests/tools/marginal_model.py        # prints, per resolution: brute-force marks, closed-form marks, missed (must be 0)
"""
import sys, numpy as np
sys.path.insert(0, __file__.rsplit('/tests/', 1)[0])
import freesasa_b200 as fs
from oracle import bindings as ob
n=100000
xyz,radii=fs.workloads.globule(n)
xyz=np.round(xyz,3); radii=np.round(radii,2)
R=radii+1.4
start,lst=ob.oracle_neighbours(xyz,R)
rng=np.random.default_rng(1)
rad=np.linalg.norm(xyz,axis=1)
cand=np.where(rad>rad.max()-8)[0]
sample=rng.choice(cand,3000,replace=False)
def run(ns, qmin):
    tot_b=tot_a=miss=0; npairs=0
    worst=[]
    for i in sample:
        nb=lst[start[i]:start[i+1]]
        if len(nb)==0: continue
        D=xyz[nb]-xyz[i]; Rj=R[nb]; Ri=R[i]
        dz=D[:,2]; d=np.hypot(D[:,0],D[:,1]); D3=np.sqrt(d*d+dz*dz)
        delta=2*Ri/ns
        zs=-Ri+(np.arange(ns)+0.5)*delta
        # brute force
        a2=(Ri-np.abs(zs))*(Ri+np.abs(zs)); a=np.sqrt(np.maximum(a2,0))[:,None]
        dj=np.abs(dz[None,:]-zs[:,None]); b2=(Rj[None,:]-dj)*(Rj[None,:]+dj); act=b2>0; b=np.sqrt(np.maximum(b2,0))
        dd=d[None,:]
        N=(a+b-dd)*(dd+b-a); Dn=(dd+a-b)*(a+b+dd)
        lo=np.minimum(np.abs(N),np.abs(Dn)); hi=np.maximum(np.abs(N),np.abs(Dn))
        near=act&(lo<qmin*hi)
        brute=near.any(1)
        # closed form, in fp32 as the kernel evaluates it (numpy float32 arithmetic; the kernel's MUFU approximations are
        # within the same slack)
        f=np.float32
        Rif=f(Ri); dzf=dz.astype(f); Rjf=Rj.astype(f); df=d.astype(f)
        deltaf=f(2)*Rif/f(ns); inv_delta=f(ns)/(f(2)*Rif)
        D3sq=df*df+dzf*dzf
        inv_D3=f(1)/np.sqrt(D3sq)
        t=f(0.5)*(D3sq+(Rif-Rjf)*(Rif+Rjf))*inv_D3
        rho2=(Rif-t)*(Rif+t)
        ok=rho2>0
        rho=np.sqrt(np.maximum(rho2,f(0)))
        zc=t*dzf*inv_D3; ext=rho*df*inv_D3
        slack=f(4e-6)*(f(1)+f(1)/np.maximum(rho,f(1e-3)))
        flags=np.zeros(ns,bool)
        zsf=(-Rif+(np.arange(ns).astype(f)+f(0.5))*deltaf).astype(f)
        for zcrit in (zc-ext, zc+ext):
            z=zcrit
            aa=np.sqrt(np.maximum((Rif-z)*(Rif+z),f(1e-12))); zz=z-dzf; bb=np.sqrt(np.maximum((Rjf-zz)*(Rjf+zz),f(0)))
            Nv=np.abs((aa+bb-df)*(df+bb-aa)); Dv=np.abs((df+aa-bb)*(aa+bb+df))
            zda=f(2)*z*df/aa
            slope=np.where(Nv<Dv,np.abs(f(2)*dzf-zda),np.abs(-f(2)*dzf-zda)); other=np.maximum(Nv,Dv)
            eps=np.minimum(np.maximum(f(2)*f(qmin)*other/np.maximum(slope,f(1e-9)),f(2e-6))+f(2e-6)+slack,deltaf)
            sidx=np.rint((z+Rif)*inv_delta-f(0.5)).astype(int)
            for k in (-1,0,1):
                s_=sidx+k
                v=ok&(s_>=0)&(s_<ns)
                s2=np.clip(s_,0,ns-1)
                hit=v&(np.abs(zsf[s2]-z)<eps)
                flags[s2[hit]]=True
        close=ok&(f(2)*ext<f(1e-3))
        for j in np.where(close)[0]:
            w=f(1e-4)+slack[j]
            flags|=(zs>zc[j]-ext[j]-w)&(zs<zc[j]+ext[j]+w)
        tot_b+=brute.sum(); tot_a+=flags.sum(); m=(brute&~flags); miss+=m.sum(); npairs+=len(nb)*ns
        if m.any():
            for s_ in np.where(m)[0]:
                j=np.where(near[s_])[0][0]
                worst.append((i,s_,j,lo[s_,j]/hi[s_,j], zs[s_]-(zc[j]-ext[j]), zs[s_]-(zc[j]+ext[j]), d[j], dz[j], rho[j]))
    print(f"ns={ns} qmin={qmin:g}: brute flagged {tot_b}, analytic flagged {tot_a}, missed {miss} (pairs-slices {npairs})")
    for w in worst[:8]: print("   miss:", w)
for ns,q in ((100,3e-6),(20,3e-6),(5,2.4e-5)):
    run(ns,q)
