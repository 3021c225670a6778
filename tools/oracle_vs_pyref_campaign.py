#!/usr/bin/env python
"""Random campaigns of the C++ oracle against the independent Python / numpy transcriptions of tests/pyref.py (CPU only):

    python tools/oracle_vs_pyref_campaign.py twisted SEED SECONDS     # FIND_NEAREST_POINT on ORCA-shaped sources
    python tools/oracle_vs_pyref_campaign.py akima   SEED SECONDS     # AKIMA_2D on regular sources
    python tools/oracle_vs_pyref_campaign.py drown   SEED SECONDS     # BDROWN / DROWN / SMOOTHER on random masks and options
    python tools/oracle_vs_pyref_campaign.py vert    SEED SECONDS     # AKIMA_1D_3D / AKIMA_1D_1D with missing values and persistence options
    python tools/oracle_vs_pyref_campaign.py mapping SEED SECONDS     # MAPPING_BL (search by the oracle, body per target) on both source kinds

Test infrastructure: nothing here is part of the product."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ["FUZZ_ORACLE_ONLY"] = "1"
import numpy as np

from oracle import oracle as orc
from sosie_b200 import grids as G
from tests import pyref as P
import fuzz_parity as F


def twisted(seed, seconds):
    rng=np.random.default_rng(seed)
    t0=time.time(); n=0; bad=0
    while time.time()-t0 < seconds:
        piv=str(rng.choice(["F","T"]))
        Xs,Ys=G.orca_like(int(rng.integers(24,60)),int(rng.integers(24,44)),piv,float(rng.uniform(0,200)))
        Xt,Yt=F.rnd_target(rng,Xs,Ys)
        if np.ndim(Xt)==1:
            Xt,Yt=G.expand_2d(Xt,Yt)
        if Xt.size>1500: continue
        mask=None
        if rng.random()<0.3: mask=(rng.random(Xt.shape)>0.2).astype(np.int32)
        orc.set_libm(orc.GLIBC)
        try:
            JIo,JJo,mo,used,info=orc.find_nearest_point(Xt,Yt,Xs,Ys,mask=mask)
            err=orc.get_error()
        finally:
            orc.set_libm(orc.PMATH)
        XsF,YsF,XtF,YtF=P.to_f(Xs),P.to_f(Ys),P.to_f(Xt),P.to_f(Yt)
        try:
            reg_s,reg_t,js,je,icr=P.find_nearest_prelude(XtF,YtF,XsF,YsF)
            if reg_s: continue
            m0=np.ones(XtF.shape,np.int32) if mask is None else P.to_f(mask).astype(np.int32)
            if js<1 or je<1 or js>XtF.shape[1] or je>XtF.shape[1]:
                perr=True
            else:
                JI,JJ,m=P.find_nearest_twisted(XtF,YtF,XsF,YsF,m0,js,je,icr); perr=False
        except RuntimeError as e:
            perr=True
        n+=1
        if perr or err:
            if bool(perr)!=bool(err):
                bad+=1; print("STOP mismatch", perr, err, Xs.shape, Xt.shape, flush=True)
            continue
        ok=(tuple(int(v) for v in info[:3])==(js,je,icr)) and np.array_equal(P.from_f(JI),JIo) and np.array_equal(P.from_f(JJ),JJo) and np.array_equal(P.from_f(m),mo)
        if not ok:
            bad+=1; print("MISMATCH", Xs.shape, Xt.shape, info, (js,je,icr), (P.from_f(JI)!=JIo).sum(), flush=True)
    print("cases",n,"bad",bad)


def akima(seed, seconds):
    rng=np.random.default_rng(seed)
    t0=time.time(); n=0; bad=0
    while time.time()-t0 < seconds:
        lon_s,lat_s,kew=F.rnd_regular_axes(rng)
        if lon_s.size<6 or lat_s.size<6 or np.any(np.diff(lon_s)<=0): continue
        X2,Y2=G.expand_2d(lon_s,lat_s)
        Xt,Yt=F.rnd_target(rng,X2,Y2)
        if np.ndim(Xt)==1: Xt,Yt=G.expand_2d(Xt,Yt)
        if Xt.size>1200: continue
        Z1=G.field_slabs(X2,Y2,1,seed=int(rng.integers(1<<30)))
        if rng.random()<0.3: Z1[0,2:6,3:9]=1.5
        with np.errstate(all="ignore"):
            Z2o,ixo=orc.akima_2d(kew,X2,Y2,Z1,Xt,Yt)
            Z2p,ixp=P.akima_2d(kew,P.to_f(X2),P.to_f(Y2),P.to_f(Z1[0]),P.to_f(Xt),P.to_f(Yt))
        Z2p=P.from_f(Z2p); n+=1
        ok=np.array_equal(ixo[0],P.from_f(ixp[:,:,0])) and np.array_equal(ixo[1],P.from_f(ixp[:,:,1])) and np.array_equal(np.isnan(Z2o[0]),np.isnan(Z2p)) and np.array_equal(Z2o[0][~np.isnan(Z2p)],Z2p[~np.isnan(Z2p)])
        if not ok:
            bad+=1; print("MISMATCH",X2.shape,Xt.shape,kew,(Z2o[0]!=Z2p).sum(),flush=True)
    print("cases",n,"bad",bad)


def mapping(seed, seconds):
    rng=np.random.default_rng(seed)
    t0=time.time(); n=0; bad=0; npts=0
    eps=float(np.float32(1e-9))
    while time.time()-t0 < seconds:
        curv = rng.random()<0.5
        if curv:
            Xs,Ys=G.orca_like(int(rng.integers(24,50)),int(rng.integers(24,40)),str(rng.choice(["F","T"])),float(rng.uniform(0,200))); kew=2
        else:
            lon_s,lat_s,kew=F.rnd_regular_axes(rng)
            if lon_s.size<6 or lat_s.size<6: continue
            Xs,Ys=G.expand_2d(lon_s,lat_s)
        Xt,Yt=F.rnd_target(rng,Xs,Ys)
        if np.ndim(Xt)==1: Xt,Yt=G.expand_2d(Xt,Yt)
        if Xt.size>400: continue
        if not curv and rng.random()<0.5:   # snap some targets onto source nodes / grid lines
            k=rng.integers(0,Xt.size,8); Xt.flat[k]=Xs.flat[rng.integers(0,Xs.size,8)] % 360.0
            k=rng.integers(0,Xt.size,8); Yt.flat[k]=np.clip(Ys.flat[rng.integers(0,Ys.size,8)],-89.9,89.9)
        orc.set_libm(orc.GLIBC)
        try:
            m=orc.mapping_bl(kew,Xs,Ys,Xt,Yt); err=orc.get_error()
        finally:
            orc.set_libm(orc.PMATH)
        if err: continue
        met,ab,ipb=m["metrics"],m["alphabeta"],m["ipb"]
        XsF,YsF=P.to_f(Xs),P.to_f(Ys)
        n+=1
        for j in range(Xt.shape[0]):
            for i in range(Xt.shape[1]):
                iP,jP=int(met[0,j,i]),int(met[1,j,i])
                if iP<1: continue
                try:
                    r=P.mapping_point_2d(kew,XsF,YsF,Xt[j,i],Yt[j,i],iP,jP)
                except Exception as e:
                    r="EXC"
                if r=="EXC": continue
                if r is None:
                    ok = met[2,j,i]==1 and ipb[j,i]==1 and ab[0,j,i]==0.0 and ab[1,j,i]==0.0
                else:
                    iq,a,b,ip=r
                    if a<0 and a>-eps: a=0.0
                    if b<0 and b>-eps: b=0.0
                    if a>-9999.0 and a<0: a,ip=0.5,4
                    if a>1: a,ip=0.5,5
                    if b>-9999.0 and b<0: b,ip=0.5,6
                    if b>1: b,ip=0.5,7
                    if iq<1: iq,ip=1,1
                    same=lambda u,v: (u==v) or (u!=u and v!=v)
                    ok=int(met[2,j,i])==iq and same(ab[0,j,i],a) and same(ab[1,j,i],b) and int(ipb[j,i])==ip
                npts+=1
                if not ok:
                    bad+=1
                    if bad<10: print("MISMATCH",curv,kew,Xs.shape,(i,j),(iP,jP),Xt[j,i],Yt[j,i],r,(int(met[2,j,i]),ab[0,j,i],ab[1,j,i],int(ipb[j,i])),flush=True)
    print("grids",n,"targets",npts,"bad",bad)


def drown(seed, seconds):
    rng=np.random.default_rng(seed)
    t0=time.time(); n=0; bad=0
    def eq(a,b):
        return np.array_equal(np.isnan(a),np.isnan(b)) and np.array_equal(a[~np.isnan(a)],b[~np.isnan(b)])
    while time.time()-t0 < seconds:
        ni,nj=int(rng.integers(5,40)),int(rng.integers(4,30))
        kew=int(rng.choice([-1,0,2]))
        if kew>=0 and ni<2*kew+3: continue
        X=rng.normal(10.0,5.0,(nj,ni)).astype(np.float32)
        land=rng.random((nj,ni))<rng.uniform(0.0,0.9)
        if rng.random()<0.5:
            yy,xx=np.mgrid[0:nj,0:ni]
            land=(np.sin(xx*rng.uniform(0.05,0.5))*np.cos(yy*rng.uniform(0.05,0.5))>rng.uniform(-0.5,0.8))
        if rng.random()<0.05: land[...]=True
        mask=(~land).astype(np.int32)
        X[mask==0]=-9999.0
        if rng.random()<0.2: X.flat[rng.integers(0,X.size,3)]=np.float32(-0.0)
        ninc,nsm=int(rng.integers(0,25)),int(rng.integers(0,8))
        pmin,pmax=(-1.0e3,1.0e3) if rng.random()<0.7 else (5.0,12.0)
        pval=float(rng.choice([0.0,-9999.0,1.0e-13]))
        which=int(rng.integers(0,3))
        with np.errstate(all="ignore"):
            if which==0:
                a,ns=orc.bdrown(kew,X,mask,ninc,nsm,pmin,pmax,pval); b,nsp=P.bdrown(kew,P.to_f(X),P.to_f(mask),ninc,nsm,pmin,pmax,pval); b=P.from_f(b)
                ok=eq(a,b) and int(ns[0])==nsp
            elif which==1:
                m8=mask.astype(np.int8); k=min(ninc,6)
                a,ns=orc.drown(kew,X,m8,k); b,nsp=P.drown(kew,P.to_f(X),P.to_f(m8),k); b=P.from_f(b)
                ok=eq(a,b) and int(ns[0])==nsp
            else:
                use=rng.random()<0.6; lemp=bool(use and rng.random()<0.5)
                a=orc.smoother(kew,X,nsm,mask if use else None,lemp); b=P.from_f(P.smoother(kew,P.to_f(X),nsm,msk=P.to_f(mask) if use else None,l_emp=lemp))
                ok=eq(a,b)
        n+=1
        if not ok:
            bad+=1
            if bad<8: print("MISMATCH",which,ni,nj,kew,ninc,nsm,pmin,pval,(a!=b).sum(),flush=True)
    print("cases",n,"bad",bad)


def vert(seed, seconds):
    pyref = P
    rng=np.random.default_rng(seed)
    t0=time.time(); n=0; bad=0
    while time.time()-t0 < seconds:
        which=int(rng.integers(0,2))
        if which==0:
            nk1,nk2=int(rng.integers(3,20)),int(rng.integers(1,25))
            vz1=np.sort(rng.uniform(0,5500,nk1)).astype(np.float32)
            if len(np.unique(vz1))<nk1: continue
            vz2=np.sort(rng.uniform(0,6000,nk2)).astype(np.float32)
            if rng.random()<0.4 and nk2>1: vz2[int(rng.integers(0,nk2))]=vz1[int(rng.integers(0,nk1))]; vz2=np.sort(vz2)
            xf1=rng.normal(10,4,(nk1,3,4)).astype(np.float32)
            if rng.random()<0.3: xf1[int(rng.integers(0,nk1)):, :, 1]=-9999.0
            with np.errstate(all="ignore"):
                a=orc.akima_1d_3d(vz1,xf1,vz2,-9999.0); b=pyref.akima_1d_3d_np(vz1,xf1,vz2,-9999.0)
            ok=np.array_equal(a.view(np.uint32),b.view(np.uint32))
        else:
            ncol,nk1,nk2=4,int(rng.integers(3,16)),int(rng.integers(1,20))
            vz1=np.sort(rng.uniform(0,5500,(ncol,nk1)),axis=1); vz2=np.sort(rng.uniform(0,6000,(ncol,nk2)),axis=1)
            vf1=rng.normal(10,4,(ncol,nk1))
            opts={}
            if rng.random()<0.5: opts["l_surf_persist"]=True
            if rng.random()<0.5: opts["l_botm_persist"]=True
            if rng.random()<0.5:
                opts["rmask_val"]=-9999.0
                for c in range(ncol):
                    r=rng.random()
                    if r<0.3: vf1[c,int(rng.integers(1,nk1)):]=-9999.0
                    elif r<0.4: vf1[c,:int(rng.integers(1,nk1))]=-9999.0
                    elif r<0.5: vf1[c,int(rng.integers(0,nk1))]=-9999.0
            with np.errstate(all="ignore"):
                out,st=orc.akima_1d_1d(vz1,vf1,vz2,**opts)
                ok=True
                for c in range(ncol):
                    b,sb=pyref.akima_1d_1d_py(vz1[c],vf1[c],vz2[c],opts.get("rmask_val"),opts.get("l_surf_persist",False),opts.get("l_botm_persist",False))
                    if st[c]!=sb: ok=False
                    elif sb>=0 and not np.array_equal(out[c].view(np.uint64) if out.dtype==np.float64 else out[c].view(np.uint32), b.view(np.uint64) if b.dtype==np.float64 else b.view(np.uint32)): ok=False
        n+=1
        if not ok:
            bad+=1
            if bad<8: print("MISMATCH",which,nk1,nk2,opts if which else "",flush=True)
    print("cases",n,"bad",bad)


if __name__ == "__main__":
    orc.build()
    orc.set_libm(orc.PMATH)
    {"twisted": twisted, "akima": akima, "mapping": mapping, "drown": drown, "vert": vert}[sys.argv[1]](int(sys.argv[2]), float(sys.argv[3]))
