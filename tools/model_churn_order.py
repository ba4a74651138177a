#!/usr/bin/env python
"""CPU model of what streaming churn does to the Transform-pool ORDER (no GPU needed): the churn pattern of round 1 (4 % of
the instances despawn per frame as whole groups, as many spawn) replayed on the real pool mirror (scgpu_pool.h through
tests/hostsim, i.e. the reference's swap-with-last order), reporting per frame how many instances end up 32 or more slots away from their parent (a
link no hierarchy window can hold: k_update_win hands those windows to the generic path) and how many have their parent
AFTER them. Result (1 Mi instances, depth-4 groups): far links 1.2 % after one frame, 8.8 % after 8, 20 % after 20,
33 % after 40; windows (the greedy cut of k_build_windows replayed) holding at least one far
child: 11 %, 59 %, 86 %, 94 % — see DESIGN.md §9."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'sc-gameengine_b200'))
from scgpu import scenes
hs = C.CDLL(str(ROOT / 'tests' / 'hostsim' / 'libhostsim.so'))
hs.hs_pool_create.restype = C.c_void_p; hs.hs_pool_create.argtypes=[C.c_uint32]
hs.hs_pool_spawn.argtypes=[C.c_void_p,C.c_uint32,C.c_void_p,C.c_void_p]
hs.hs_pool_despawn.restype=C.c_double; hs.hs_pool_despawn.argtypes=[C.c_void_p,C.c_uint32,C.c_void_p]
hs.hs_pool_count.restype=C.c_uint32; hs.hs_pool_count.argtypes=[C.c_void_p]
hs.hs_pool_read.argtypes=[C.c_void_p]+[C.c_void_p]*4
n = 1<<20
sc = scenes.city_hier(n)
e = np.arange(n, dtype=np.uint32)
IDX = 1<<24
parent_of = np.full(IDX, -1, np.int64)      # entity index -> parent entity index
parent_of[:n] = sc["parent"]
pool = hs.hs_pool_create(IDX); hs.hs_pool_spawn(pool, n, e.ctypes.data, None)
rng = np.random.default_rng(5)
roots = np.nonzero(sc["parent"] < 0)[0]; group_end = np.append(roots[1:], n)
groups = [(int(a),int(b)) for a,b in zip(roots, group_end)]   # entity id ranges
alive = np.ones(len(groups), bool)
tmpl = scenes.city_hier(n//10+64, seed=99)
next_id = n
for f in range(40):
    cand = np.nonzero(alive)[0]
    pick = rng.choice(cand, max(1,len(cand)//25), replace=False)   # ~4 % like the bench
    alive[pick] = False
    dead = np.concatenate([np.arange(groups[g][0], groups[g][1], dtype=np.uint32) for g in pick])
    hs.hs_pool_despawn(pool, len(dead), dead.ctypes.data)
    m = len(dead)
    fresh = np.arange(next_id, next_id+m, dtype=np.uint32)
    tp = tmpl["parent"][:m]; tp = np.where(tp < m, tp, -1)
    parent_of[fresh] = np.where(tp>=0, tp+next_id, -1)
    # register fresh groups as alive groups
    tr = np.nonzero(tp<0)[0]; te = np.append(tr[1:], m)
    for a,b in zip(tr,te): groups.append((next_id+int(a), next_id+int(b)))
    alive = np.concatenate([alive, np.ones(len(tr), bool)])
    hs.hs_pool_spawn(pool, m, fresh.ctypes.data, None)
    next_id += m
    cnt = hs.hs_pool_count(pool)
    dense = np.zeros(cnt, np.uint32); hs.hs_pool_read(pool, dense.ctypes.data, None, None, None)
    slot_of = np.full(IDX, -1, np.int64); slot_of[dense] = np.arange(cnt)
    par = parent_of[dense]; has = par >= 0
    ps = np.where(has, slot_of[np.where(has, par, 0)], -1)
    dist = np.where(has, np.arange(cnt) - ps, 0)
    far = has & (np.abs(dist) > 31)      # k_build_windows: a link of 32 slots or more can never sit inside one window
    fwd = has & (dist < 0)
    # the greedy window cut of k_build_windows (largest position within 32 slots that no near link crosses) and the
    # share of windows that hold at least one far child, i.e. that k_update_win hands to the generic path
    slots = np.arange(cnt)
    near = has & ~far
    lo = np.minimum(slots, ps)[near]; hi = np.maximum(slots, ps)[near]
    diff = np.zeros(cnt + 2, np.int64)
    np.add.at(diff, lo + 1, 1); np.add.at(diff, hi + 1, -1)
    crossed = np.cumsum(diff)[: cnt + 1] > 0          # crossed[c]: a near link crosses the cut position c
    far_prefix = np.concatenate([[0], np.cumsum(far)])
    start, n_win, n_slow, inst_slow = 0, 0, 0, 0
    while start < cnt:
        end = min(start + 32, cnt)
        c = end
        while c > start + 1 and c < cnt and crossed[c]:
            c -= 1
        if c <= start:
            c = end
        n_win += 1
        if far_prefix[c] - far_prefix[start] > 0:
            n_slow += 1; inst_slow += c - start
        start = c
    if f % 4 == 3 or f < 4:
        print(f"          windows {n_win} (mean {cnt / n_win:.1f} slots), with a far child: {n_slow} ({100 * n_slow / n_win:.1f} %) holding {100 * inst_slow / cnt:.1f} % of the instances")
    if f % 4 == 3 or f < 4:
        print(f"frame {f+1:2d}: instances {cnt}, children more than 31 slots from their parent: {far.sum()} ({100*far.mean():.2f} %); parent AFTER child (any distance): {fwd.sum()} ({100*fwd.mean():.2f} %)")
