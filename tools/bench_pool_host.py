#!/usr/bin/env python
"""Host-only timing of the Transform-pool mirror (sc-gameengine_b200/csrc/scgpu_pool.h through tests/hostsim; no GPU):
an 8 Mi pool under the churn pattern of round 1 — 4 % of the instances despawn per frame as whole groups, as many spawn —
for 1, 4 and 8 host threads.  python tools/bench_pool_host.py [instances]

The numbers in DESIGN.md §8c (9.3 ms on one thread, 5.8 ms inside scgpuDespawn with four) were taken on the GPU box's
host with caches cold from the harness's own frame work; this script leaves less between the calls and reads lower."""
import ctypes as C
import statistics
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
HS = ROOT / "tests" / "hostsim"


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8 * 1024 * 1024
    subprocess.run(["make", "-C", str(HS)], check=True, capture_output=True)
    L = C.CDLL(str(HS / "libhostsim.so"))
    L.hs_pool_create.restype = C.c_void_p
    L.hs_pool_create.argtypes = [C.c_uint32]
    L.hs_pool_destroy.argtypes = [C.c_void_p]
    L.hs_pool_set_threads.argtypes = [C.c_void_p, C.c_uint32]
    L.hs_pool_spawn.restype = C.c_int
    L.hs_pool_spawn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.hs_pool_despawn.restype = C.c_double
    L.hs_pool_despawn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.hs_pool_num_removed.restype = C.c_uint32
    L.hs_pool_num_removed.argtypes = [C.c_void_p]
    group = 6
    for threads in (1, 4, 8):
        rng = np.random.default_rng(5)
        p = L.hs_pool_create(1 << 24)
        L.hs_pool_set_threads(p, threads)
        e = np.arange(n, dtype=np.uint32)
        L.hs_pool_spawn(p, n, e.ctypes.data, None)
        alive_groups = np.arange(n // group, dtype=np.int64)     # group g = entity indices [g*group, (g+1)*group)
        next_group = n // group
        ms = []
        for frame in range(10):
            pick = rng.choice(len(alive_groups), len(alive_groups) // 25, replace=False)
            dead_groups = alive_groups[pick]
            alive_groups = np.delete(alive_groups, pick)
            dead = (dead_groups[:, None] * group + np.arange(group)[None, :]).ravel().astype(np.uint32)
            sec = L.hs_pool_despawn(p, len(dead), dead.ctypes.data)
            assert L.hs_pool_num_removed(p) == len(dead)
            fresh_groups = (np.arange(len(dead_groups), dtype=np.int64) + next_group) % ((1 << 24) // group)
            next_group += len(dead_groups)
            fresh = (fresh_groups[:, None] * group + np.arange(group)[None, :]).ravel().astype(np.uint32)
            if L.hs_pool_spawn(p, len(fresh), fresh.ctypes.data, None) == 0:
                alive_groups = np.concatenate([alive_groups, fresh_groups])
            if frame >= 2:
                ms.append(sec * 1e3)
        print(f"threads {threads}: {len(dead)} despawns per frame, replay median {statistics.median(ms):.2f} ms "
              f"(min {min(ms):.2f}, max {max(ms):.2f})")
        L.hs_pool_destroy(p)


if __name__ == "__main__":
    main()
