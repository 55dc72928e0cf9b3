import sys, time
sys.path[:0]=['/root/repo','/root/repo/oracle','/root/repo/tests']
import numpy as np, common, oracle as orc
import cryogrid_jl_b200 as cg
variant=sys.argv[1]
tile = common.cfg4_tile(2, variant)
H0 = cg.initialcondition(tile)
oc = orc.OracleColumn(tile, 0)
H,t,dt=H0[:,0],0.0,60.0
rng=np.random.default_rng(0)
t0=time.time()
for d in (20,75,130,190,250,310,360):
    H,t,dt,ns,st,tr,nl = oc.advance_trace(H,t,dt,86400.0*d)
    for W in (1,3,6):
        Hr,_,dtr,n1,_,tr1,_ = oc.advance_trace(H,t,dt,t+W*3600)
        worst=[0,0,0,0]
        for rep in range(3):
            Hp = H*(1+2.2e-16*rng.integers(-3,4,H.size))
            Hq,_,dtq,n2,_,tr2,_ = oc.advance_trace(Hp,t,dt,t+W*3600)
            oc.rhs(Hr,t+W*3600); T1=oc.work.get("T"); W1=oc.work.get("thw"); oc.rhs(Hq,t+W*3600); T2=oc.work.get("T"); W2=oc.work.get("thw")
            m=min(len(tr1),len(tr2))
            worst=[max(worst[0],np.max(np.abs(T1-T2))),max(worst[1],np.max(np.abs(W1-W2))),max(worst[2],np.max(np.abs(tr1[:m]-tr2[:m])/tr1[:m])), max(worst[3], abs(n1-n2), abs(dtr-dtq)/dtr)]
        print(variant,d,W,n1,"dT %.1e dthw %.1e dtseq %.1e  nsteps/dtfinal %.1e"%tuple(worst), "limfrac %.2f"%(nl/ns), flush=True)
print(time.time()-t0)
