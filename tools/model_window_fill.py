#!/usr/bin/env python
"""CPU model of how full the hierarchy windows of k_update_win are under sector churn (no GPU): plays the host half of
scgpuSpawn / scgpuDespawn (tests/hostsim: pool mirror + slot layout with its window packing) on a city of vehicles (10
slots) and peds (4 slots) in random order, lets a tenth of the sectors stream out and in again per frame in a fresh
random order, and cuts the slots into windows like k_build_windows does (a window takes whole groups while they fit in
32 slots; a hole is a group of one).  python tools/model_window_fill.py [sectors groups_per_sector]

DESIGN.md section 3 quotes its output: 31.5 live slots per window after the initial spawn, 30.8 after 8 frames and 29.7
after 40 with 180-slot sectors; group-by-group placement in arrival order (the layout before the packing): 28.7 -> 28.4."""
import ctypes, subprocess, sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
subprocess.run(['make', '-C', str(ROOT / 'tests' / 'hostsim')], check=True, capture_output=True)
L=ctypes.CDLL(str(ROOT / 'tests' / 'hostsim' / 'libhostsim.so'))
L.hs_scene_create.restype=ctypes.c_void_p
L.hs_scene_create.argtypes=[ctypes.c_uint32,ctypes.c_uint32]
L.hs_scene_spawn.argtypes=[ctypes.c_void_p,ctypes.c_uint32,ctypes.c_void_p,ctypes.c_void_p,ctypes.c_void_p,ctypes.c_void_p]
L.hs_scene_despawn.argtypes=[ctypes.c_void_p,ctypes.c_uint32,ctypes.c_void_p]
L.hs_scene_extent.argtypes=[ctypes.c_void_p]; L.hs_scene_free.argtypes=[ctypes.c_void_p]
rng=np.random.default_rng(3)
gidc=[0]
def mk(G, first):
    sizes=np.where(rng.random(G)<0.5,10,4)
    n=int(sizes.sum())
    starts=np.r_[0,np.cumsum(sizes)[:-1]]
    e=np.arange(first,first+n,dtype=np.uint32)
    par=np.full(n,0xFFFFFFFF,np.uint32)
    idx=np.arange(n); gi=np.repeat(np.arange(G),sizes); off=idx-starts[gi]
    veh=sizes[gi]==10
    pk=np.where(veh, np.select([off==0,off==1,off<=5],[-1,0,1],default=off-4), off-1)
    m=pk>=0
    par[m]=e[(starts[gi]+pk)[m]]
    g=gi+gidc[0]; gidc[0]+=G
    return e,par,g
def fill(group_of_slot):
    n=len(group_of_slot)
    bounds=np.flatnonzero(np.diff(group_of_slot)!=0)+1
    bounds=np.concatenate([bounds,[n]])
    windows=0; start=0
    while start<n:
        j=np.searchsorted(bounds,start+32,side='right')-1
        nxt=bounds[j] if j>=0 and bounds[j]>start else start+32
        start=int(nxt); windows+=1
    return windows
NS = int(sys.argv[1]) if len(sys.argv) > 2 else 6000
GPS = int(sys.argv[2]) if len(sys.argv) > 2 else 26   # sectors, groups per sector (26 ~ 180 slots, the bench's city)
cap=NS*GPS*8+200000
h=L.hs_scene_create(cap, 1<<24)
gos=np.full(cap,-1,np.int64)   # group of slot; holes: unique negative ids
slot_of={}
nxt=1; r0=ctypes.c_uint32(); sect={}
es=[];ps=[];gs=[]
for s in range(NS):
    e,par,g=mk(GPS,nxt); nxt+=len(e); sect[s]=e; es.append(e); ps.append(par); gs.append(g)
e=np.concatenate(es); par=np.concatenate(ps); g=np.concatenate(gs); slot=np.zeros(len(e),np.uint32)
L.hs_scene_spawn(h,len(e),e.ctypes.data,par.ctypes.data,slot.ctypes.data,ctypes.byref(r0))
gos[slot]=g
eslot=np.zeros(1<<24,np.uint32); eslot[e&0xFFFFFF]=slot
def report(tag):
    ext=L.hs_scene_extent(h)
    a=gos[:ext].copy()
    holes=np.flatnonzero(a<0); a[holes]=-1-holes   # every hole slot its own group
    live=ext-len(holes)
    w=fill(a)
    print(tag,'extent',ext,'live',live,'windows',w,'live/window',round(live/w,2))
report('initial')
for it in range(40):
    ids=rng.choice(NS,NS//10,replace=False)
    vict=np.concatenate([sect[s] for s in ids])
    L.hs_scene_despawn(h,len(vict),vict.ctypes.data)
    gos[eslot[vict&0xFFFFFF]]=-1
    es=[];ps=[];gs=[]
    for s in rng.permutation(ids):
        e2,par2,g2=mk(GPS,nxt); nxt+=len(e2); sect[s]=e2; es.append(e2); ps.append(par2); gs.append(g2)
    e2=np.concatenate(es); par2=np.concatenate(ps); g2=np.concatenate(gs); s2=np.zeros(len(e2),np.uint32)
    L.hs_scene_spawn(h,len(e2),e2.ctypes.data,par2.ctypes.data,s2.ctypes.data,ctypes.byref(r0))
    gos[s2]=g2; eslot[e2&0xFFFFFF]=s2
    if it in (0,1,3,7,15,39): report('frame %d'%it)
