import sys, time
sys.path[:0]=['/root/repo','/root/repo/oracle','/root/repo/tests']
import numpy as np, common, oracle as orc
import cryogrid_jl_b200 as cg
tile = common.samoylov_freew_tile(ncolumns=2, seed=1)
H0 = cg.initialcondition(tile)
oc = orc.OracleColumn(tile, 0)
H,t,dt=H0[:,0],0.0,60.0
rng=np.random.default_rng(0)
for d in (20,75,130,190,250,310,360):
    H,t,dt,ns,st,tr,nl = oc.advance_trace(H,t,dt,86400.0*d)
    for W in (1,3,6,24):
        Hr,_,_,n1,_,tr1,_ = oc.advance_trace(H,t,dt,t+W*3600)
        Hp = H*(1+2.2e-16*rng.integers(-3,4,H.size))
        Hq,_,_,n2,_,tr2,_ = oc.advance_trace(Hp,t,dt,t+W*3600)
        oc.rhs(Hr,t+W*3600); T1=oc.work.get("T"); oc.rhs(Hq,t+W*3600); T2=oc.work.get("T")
        m=min(len(tr1),len(tr2))
        print(d,W,n1,n2,"dT",np.max(np.abs(T1-T2)),"dt rel",np.max(np.abs(tr1[:m]-tr2[:m])/tr1[:m]), "nlim", nl/ns)
