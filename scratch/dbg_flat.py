import sys, numpy as np
sys.path.insert(0,'sc-gameengine_b200'); sys.path.insert(0,'tests')
import scgpu
from scgpu import scenes
n=1_000_000; views=int(sys.argv[1]) if len(sys.argv)>1 else 1
sc=scenes.city_flat(n); e=np.arange(n,dtype=np.uint32)
par=scenes.parent_handles(sc["parent"],e); vps=scenes.standard_views(views)
for trial in range(4):
    s=scgpu.Scene(n,max_views=views,max_entity_index=n)
    s.spawn(e,sc["trs9"],par,sc["aabb6"],sc["mesh_mat"],sc["flags"]); s.set_views(vps)
    s.update(); w1=s.read_world(e).reshape(n,16).copy()
    s.mark_all_dirty(); s.update(); w2=s.read_world(e).reshape(n,16).copy()
    s.mark_all_dirty(); s.update(); w3=s.read_world(e).reshape(n,16).copy()
    for name,(a,b) in {"1v2":(w1,w2),"2v3":(w2,w3)}.items():
        d=np.argwhere(a.view(np.uint32)!=b.view(np.uint32))
        rows=np.unique(d[:,0])
        print(trial,name,"rows differing",len(rows),[(int(r),int(r)%1024,int(r)%32,sorted(set(d[d[:,0]==r,1].tolist()))) for r in rows[:12]])
    s.close()
