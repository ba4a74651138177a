"""Stress test used to find the cross-proxy WAR race of the TMA-staged flat kernel: repeats spawn -> update -> read back of a 1M-instance flat city and reports any frame-to-frame difference (python tools/stress_flat_determinism.py 100)."""
import sys, numpy as np
sys.path.insert(0,'sc-gameengine_b200'); sys.path.insert(0,'tests')
import scgpu
from scgpu import scenes
n=1_000_000; views=1
sc=scenes.city_flat(n); e=np.arange(n,dtype=np.uint32)
par=scenes.parent_handles(sc["parent"],e); vps=scenes.standard_views(views)
sample=np.random.default_rng(1).choice(n,50_000,replace=False).astype(np.uint32)
ref=None
for trial in range(int(sys.argv[1])):
    s=scgpu.Scene(n,max_views=views,max_entity_index=n)
    s.spawn(e,sc["trs9"],par,sc["aabb6"],sc["mesh_mat"],sc["flags"]); s.set_views(vps)
    s.update(); c1=s.counts(); l1=s.read_visible(0)
    w1=s.read_world(sample).reshape(-1,16).copy()
    wf=s.read_world(e).reshape(-1,16).copy()
    s.mark_all_dirty(); s.update(); w2=s.read_world(sample).reshape(-1,16).copy()
    if ref is None: ref=w2
    for name,(a,b) in {"w1vref":(w1,ref),"w2vref":(w2,ref),"wfull_v_ref":(wf[sample],ref)}.items():
        d=np.argwhere(a.view(np.uint32)!=b.view(np.uint32))
        rows=np.unique(d[:,0])
        if len(rows): print(trial,name,"rows differing",len(rows),[(int(sample[r]),int(sample[r])%1024,int(sample[r])%32,sorted(set(d[d[:,0]==r,1].tolist())),a[r].tolist(),b[r].tolist()) for r in rows[:4]],flush=True)
    s.close()
print("done")
