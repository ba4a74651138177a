"""Diagnostic (python tools/probe_visible_cost.py on a B200): what the warps near a frustum cost in the fused kernel — all
dirty / 30 % dirty / nothing dirty, standard views against views that see nothing — before and after half of the groups
were despawned and respawned. It is what showed that the price of churn in round 2 was no longer broken hierarchy windows
but lost SPATIAL coherence of neighbouring slots (instruction count of the visible region x 9 with LIFO hole reuse),
which the next-fit slot layout (csrc/scgpu_layout.h) then removed."""
import sys, numpy as np
sys.path.insert(0,'sc-gameengine_b200'); sys.path.insert(0,'tests')
import scgpu
from scgpu import scenes
n=8*1024*1024; views=5
sc=scenes.city_hier(n,seed=99); e=np.arange(n,dtype=np.uint32)
s=scgpu.Scene(n+n//8,max_views=views,max_entity_index=n)
s.spawn(e,sc["trs9"],scenes.parent_handles(sc["parent"],e),sc["aabb6"],sc["mesh_mat"],sc["flags"])
std=scenes.standard_views(views); away=scenes.standard_views(views,center=(1e6,6.0,1e6))
rng=np.random.default_rng(1)
idx=np.sort(rng.choice(n,(3*n)//10,replace=False)).astype(np.uint32); trs=sc["trs9"][idx].copy()
s.enable_timings(True)
def run(vps,label,dirty):
    s.set_views(vps)
    ks=[]
    for f in range(8):
        if dirty=="all": s.mark_all_dirty()
        elif dirty=="30": s.set_local(idx,trs)
        s.update(0); c=s.counts(); k,u=s.last_timings(); ks.append(k)
    print(label, dirty, "kernel ms %.4f"%np.median(ks[2:]), "visible",[int(c.visible[v]) for v in range(views)], "recomputed",c.recomputed, flush=True)
for d in ("all","30","none"):
    run(std,"std ",d); run(away,"away",d)
# scramble ranks: despawn/respawn random groups
roots=np.nonzero(sc["parent"]<0)[0]; glen=np.diff(np.append(roots,n)); co=rng.integers(0,10,size=len(roots)); cohort=np.repeat(co,glen)
for c in range(5):
    ix=np.nonzero(cohort==c)[0].astype(np.uint32)
    s.despawn(ix)
    pos=np.full(n,-1,np.int64); pos[ix]=np.arange(len(ix))
    lp=np.where(sc["parent"][ix]>=0,pos[np.maximum(sc["parent"][ix],0)],-1)
    fresh=(ix|np.uint32(1<<24)).astype(np.uint32)
    s.spawn(fresh,sc["trs9"][ix],scenes.parent_handles(lp,fresh),sc["aabb6"][ix],sc["mesh_mat"][ix],sc["flags"][ix])
    s.update(0)
print("after churn")
for d in ("all","none"):
    run(std,"std ",d); run(away,"away",d)
