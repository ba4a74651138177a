"""Device layout decoupled from pool order (sc-gameengine_b200/csrc/scgpu_layout.h + scgpu_pool.h, the host halves of
scgpuSpawn / scgpuDespawn) checked on the CPU. The device half — k_spawn's rank / perm writes, k_despawn_apply's moves
in rank space, k_compact's rank -> perm -> entity walk — is played in numpy with the same rules:

  * the list read back in RANK order equals a naive replay of ComponentPool::add / ::remove (sc_ecs.h:203-262), i.e.
    the reference's pool order, after hundreds of frames of group-wise churn;
  * a hierarchy group spawned together stays contiguous in SLOT space for as long as it lives (what keeps the
    windows of k_update_win intact), holes of whole-group despawns are reused, the extent stays bounded;
  * no slot is ever handed out twice, live + holes == extent, a full pool still accepts element-wise placement."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

HS_DIR = Path(__file__).resolve().parent / "hostsim"
INVALID = 0xFFFFFFFF


@pytest.fixture(scope="module")
def hs():
    subprocess.run(["make", "-C", str(HS_DIR)], check=True, capture_output=True)
    L = C.CDLL(str(HS_DIR / "libhostsim.so"))
    L.hs_scene_create.restype = C.c_void_p
    L.hs_scene_create.argtypes = [C.c_uint32, C.c_uint32]
    L.hs_scene_destroy.argtypes = [C.c_void_p]
    for f in ("hs_scene_count", "hs_scene_extent", "hs_scene_free", "hs_scene_num_moves", "hs_scene_num_holes"):
        getattr(L, f).restype = C.c_uint32
        getattr(L, f).argtypes = [C.c_void_p]
    L.hs_scene_spawn.restype = C.c_int
    L.hs_scene_spawn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hs_scene_despawn.restype = C.c_uint32
    L.hs_scene_despawn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.hs_scene_read.argtypes = [C.c_void_p] * 4
    L.hs_scene_holes.argtypes = [C.c_void_p, C.c_void_p]
    return L


class NaivePool:
    """ComponentPool<Transform> as the reference writes it: add appends, remove swaps the last element in"""

    def __init__(self):
        self.dense, self.sparse = [], {}

    def add(self, e):
        self.sparse[e & 0xFFFFFF] = len(self.dense)
        self.dense.append(e)

    def remove(self, e):
        idx = e & 0xFFFFFF
        s = self.sparse.get(idx)
        if s is None or self.dense[s] != e:
            return
        last = len(self.dense) - 1
        if s != last:
            self.dense[s] = self.dense[last]
            self.sparse[self.dense[s] & 0xFFFFFF] = s
        self.dense.pop()
        del self.sparse[idx]


class DeviceModel:
    """the arrays the kernels keep: entity[slot], rank[slot], perm[rank], parent slot of every slot"""

    def __init__(self, L, capacity, sparse_size):
        self.L = L
        self.h = L.hs_scene_create(capacity, sparse_size)
        self.entity = np.full(capacity, INVALID, np.uint32)
        self.rank = np.zeros(capacity, np.uint32)
        self.perm = np.zeros(capacity, np.uint32)
        self.group = np.full(capacity, -1, np.int64)  # test bookkeeping: id of the spawn group living in the slot

    def close(self):
        self.L.hs_scene_destroy(self.h)

    @property
    def count(self):
        return self.L.hs_scene_count(self.h)

    @property
    def extent(self):
        return self.L.hs_scene_extent(self.h)

    def spawn(self, e, parent, group_ids):
        e = np.ascontiguousarray(e, np.uint32)
        parent = None if parent is None else np.ascontiguousarray(parent, np.uint32)
        slot = np.zeros(len(e), np.uint32)
        rank0 = C.c_uint32(0)
        rc = self.L.hs_scene_spawn(self.h, len(e), e.ctypes.data, None if parent is None else parent.ctypes.data,
                                   slot.ctypes.data, C.byref(rank0))
        if rc:
            return rc, None
        assert np.all(self.entity[slot] == INVALID), "a live slot was handed out again"
        assert len(np.unique(slot)) == len(slot)
        # k_spawn
        self.entity[slot] = e
        self.rank[slot] = rank0.value + np.arange(len(e), dtype=np.uint32)
        self.perm[rank0.value + np.arange(len(e))] = slot
        self.group[slot] = group_ids
        return 0, slot

    def despawn(self, e):
        e = np.ascontiguousarray(e, np.uint32)
        k = self.L.hs_scene_despawn(self.h, len(e), e.ctypes.data)
        moves = np.zeros((self.L.hs_scene_num_moves(self.h), 2), np.uint32)
        removed = np.zeros(k, np.uint32)
        self.L.hs_scene_read(self.h, None, moves.ctypes.data, removed.ctypes.data)
        # k_despawn_apply: all sources are read before any destination is written (they never overlap)
        if len(moves):
            assert moves[:, 1].min() >= self.count and moves[:, 0].max() < self.count
            s = self.perm[moves[:, 1]]
            self.perm[moves[:, 0]] = s
            self.rank[s] = moves[:, 0]
        self.entity[removed] = INVALID
        self.group[removed] = -1
        return k

    def dense_in_rank_order(self):
        """what k_compact / k_gather_dense read: entity[perm[rank]]"""
        return self.entity[self.perm[: self.count]]

    def host_dense(self):
        out = np.zeros(self.count, np.uint32)
        self.L.hs_scene_read(self.h, out.ctypes.data, None, None)
        return out

    def check_invariants(self):
        n, ext = self.count, self.extent
        live = np.flatnonzero(self.entity[:ext] != INVALID)
        assert len(live) == n, (len(live), n)
        assert np.all(self.entity[ext:] == INVALID)
        assert ext - n == self.L.hs_scene_free(self.h)
        # rank and perm are inverse permutations over the live set
        assert np.array_equal(np.sort(self.rank[live]), np.arange(n, dtype=np.uint32))
        assert np.array_equal(self.perm[self.rank[live]], live.astype(np.uint32))
        # the holes are exactly the dead slots below the extent
        holes = np.zeros((self.L.hs_scene_num_holes(self.h), 2), np.uint32)
        self.L.hs_scene_holes(self.h, holes.ctypes.data)
        dead = np.zeros(ext, bool)
        for s, ln in holes:
            assert not dead[s:s + ln].any(), "overlapping holes"
            dead[s:s + ln] = True
        assert np.array_equal(np.flatnonzero(dead), np.flatnonzero(self.entity[:ext] == INVALID))


def _mk_groups(rng, n_groups, next_index, gen, sizes=(1, 5, 10)):
    """entity handles + parent handles of n_groups hierarchy groups (root first, parents before children)"""
    e, par, gid = [], [], []
    for g in range(n_groups):
        k = int(rng.choice(sizes))
        hs_ = [(gen << 24) | (next_index + i) for i in range(k)]
        next_index += k
        for i, h in enumerate(hs_):
            e.append(h)
            par.append(INVALID if i == 0 else hs_[int(rng.integers(0, i))])
            gid.append(g)
    return np.array(e, np.uint32), np.array(par, np.uint32), np.array(gid), next_index


def test_pool_order_survives_group_churn(hs):
    rng = np.random.default_rng(7)
    cap = 40_000
    d = DeviceModel(hs, cap, 1 << 20)
    naive = NaivePool()
    nxt, gcount = 1, 0
    e, par, gid, nxt = _mk_groups(rng, 3000, nxt, 0)
    rc, slot = d.spawn(e, par, gid + gcount)
    # no holes yet: the batch takes exactly the next len(e) slots (a permutation of them: groups are packed into windows)
    assert rc == 0 and np.array_equal(np.sort(slot), np.arange(len(e)))
    gcount += 3000
    for h in e:
        naive.add(int(h))
    groups = {}  # group id -> handles
    for h, g in zip(e, gid):
        groups.setdefault(int(g), []).append(int(h))
    peak_extent = d.extent
    for frame in range(120):
        # ~10 % of the groups die as a whole (handles in group order), then as many groups spawn
        victims = rng.choice(list(groups.keys()), max(1, len(groups) // 10), replace=False)
        dead = [h for g in victims for h in groups.pop(int(g))]
        if frame % 7 == 3:
            dead += [dead[0], 0x00ABCDEF, INVALID]  # repeated / unknown / invalid handles are skipped
        d.despawn(np.array(dead, np.uint32))
        for h in dead:
            naive.remove(h)
        e, par, gid, nxt = _mk_groups(rng, len(victims), nxt, 1 + frame % 200)
        rc, slot = d.spawn(e, par, gid + gcount)
        assert rc == 0
        for h, g in zip(e, gid):
            naive.add(int(h))
            groups.setdefault(int(g) + gcount, []).append(int(h))
        gcount += len(victims)
        # the order k_compact emits == the reference's pool order
        assert np.array_equal(d.dense_in_rank_order(), np.array(naive.dense, np.uint32)), f"frame {frame}"
        assert np.array_equal(d.host_dense(), np.array(naive.dense, np.uint32))
        d.check_invariants()
        # every group still lives in consecutive slots, in spawn order
        for g, hs_ in list(groups.items())[:: max(1, len(groups) // 200)]:
            sl = np.flatnonzero(d.group == g)
            assert len(sl) == len(hs_) and sl[-1] - sl[0] == len(hs_) - 1
            assert np.array_equal(d.entity[sl], np.array(hs_, np.uint32))
        peak_extent = max(peak_extent, d.extent)
    # holes are reused: the extent stays within a few percent of the live count
    assert peak_extent <= int(d.count * 1.25) + 64, (peak_extent, d.count)
    d.close()


def test_single_member_despawns_and_full_pool(hs):
    rng = np.random.default_rng(11)
    cap = 2000
    d = DeviceModel(hs, cap, 1 << 16)
    naive = NaivePool()
    nxt = 1
    e, par, gid, nxt = _mk_groups(rng, 10_000, nxt, 0, sizes=(5,))
    e, par, gid = e[:cap], par[:cap], gid[:cap]
    assert d.spawn(e, par, gid)[0] == 0
    for h in e:
        naive.add(int(h))
    assert d.spawn(np.array([0x1000000 | 60_000], np.uint32), None, np.array([0]))[0] != 0  # pool is full
    # scattered single despawns: only 1-slot holes exist afterwards
    dead = e[rng.choice(cap, 400, replace=False)]
    dead = dead[np.argsort(rng.random(len(dead)))]
    d.despawn(dead)
    for h in dead:
        naive.remove(int(h))
    d.check_invariants()
    # 80 groups of 5 must still fit: nothing contiguous is left, so they are placed element by element
    e2, par2, gid2, nxt = _mk_groups(rng, 80, 30_000, 1, sizes=(5,))
    rc, slot = d.spawn(e2, par2, gid2 + 100_000)
    assert rc == 0 and d.count == cap and d.extent == cap
    for h in e2:
        naive.add(int(h))
    assert np.array_equal(d.dense_in_rank_order(), np.array(naive.dense, np.uint32))
    d.check_invariants()
    # a batch that does not fit is refused as a whole and changes nothing
    before = d.dense_in_rank_order().copy()
    assert d.spawn(np.array([0x2000000 | 61_000], np.uint32), None, np.array([0]))[0] != 0
    assert np.array_equal(d.dense_in_rank_order(), before)
    d.check_invariants()
    d.close()


def test_tail_despawn_shrinks_the_extent(hs):
    d = DeviceModel(hs, 1000, 1 << 12)
    e = np.arange(1, 601, dtype=np.uint32)
    assert d.spawn(e, None, np.zeros(600, np.int64))[0] == 0
    d.despawn(e[500:])          # one ascending run that ends at the extent
    assert d.extent == 500 and d.L.hs_scene_free(d.h) == 0
    d.despawn(e[100:200])       # a hole in the middle
    assert d.extent == 500 and d.L.hs_scene_free(d.h) == 100
    # a 40-member chain is cut into window-sized groups; the big hole serves them front to back
    ch = np.arange(2000, 2040, dtype=np.uint32)
    par = np.concatenate([[INVALID], ch[:-1]]).astype(np.uint32)
    rc, slot = d.spawn(ch, par, np.zeros(40, np.int64))
    assert rc == 0 and np.array_equal(slot, np.arange(100, 140))
    d.check_invariants()
    d.close()


def _window_fill(group_of_slot):
    """the greedy cut of k_build_windows over consecutive slots: a window takes whole groups while they fit in 32 slots"""
    bounds = np.flatnonzero(np.diff(group_of_slot) != 0) + 1  # positions no parent link crosses
    bounds = np.concatenate([bounds, [len(group_of_slot)]])
    windows, start = 0, 0
    i = 0
    while start < len(group_of_slot):
        j = np.searchsorted(bounds, start + 32, side="right") - 1
        nxt = bounds[j] if j >= 0 and bounds[j] > start else start + 32
        start = int(nxt)
        windows += 1
    return len(group_of_slot) / windows


def test_groups_are_packed_into_full_windows(hs):
    """vehicles of 10 and peds of 4 in random order: arrival order fills a window to 28.6 of 32 slots, the packed
    layout to > 31; nobody is placed far from its neighbours in the batch; flat batches keep their order"""
    rng = np.random.default_rng(11)
    d = DeviceModel(hs, 400_000, 1 << 20)
    e, par, gid, nxt = _mk_groups(rng, 40_000, 1, 0, sizes=(10, 4))
    rc, slot = d.spawn(e, par, gid)
    assert rc == 0 and np.array_equal(np.sort(slot), np.arange(len(e)))
    d.check_invariants()
    by_slot = np.empty(len(e), np.int64)
    by_slot[slot] = gid
    arrival = _window_fill(gid)
    packed = _window_fill(by_slot)
    assert arrival < 29.0 and packed > 31.0, (arrival, packed)
    # every group in consecutive slots, members in spawn order
    first = np.flatnonzero(np.r_[True, np.diff(gid) != 0])
    size = np.diff(np.r_[first, len(e)])
    assert np.array_equal(slot, np.repeat(slot[first], size) + (np.arange(len(e)) - np.repeat(first, size)))
    # bounded look-ahead (128 slots' worth of waiting groups): a group lands close to where arrival order would have put it
    assert np.abs(slot[first].astype(np.int64) - first).max() < 512
    # the pool order (rank) is still arrival order
    assert np.array_equal(d.dense_in_rank_order(), e)
    # a flat batch behind it: one run of fresh slots, in order
    e2 = np.arange(nxt, nxt + 5000, dtype=np.uint32)
    rc, slot2 = d.spawn(e2, np.full(5000, INVALID, np.uint32), np.arange(5000) + 100_000)
    assert rc == 0 and np.array_equal(slot2, len(e) + np.arange(5000))


def test_packing_random_group_sizes_and_nearly_full_pool(hs):
    """any mix of group sizes 1..32 (chains longer than a window are cut), churn of single groups and of long runs, a pool
    that is filled to the last slot: every element gets its own slot, groups stay consecutive, the pool order is the
    reference's, and placement always terminates"""
    rng = np.random.default_rng(23)
    cap = 6000
    d = DeviceModel(hs, cap, 1 << 18)
    naive = NaivePool()
    nxt, gcount = 1, 0
    groups = {}

    def spawn(n_groups, sizes, gen):
        nonlocal nxt, gcount
        room = cap - d.count
        e, par, gid = [], [], []
        for g in range(n_groups):
            k = int(rng.choice(sizes))
            if len(e) + k > room:
                break
            hs_ = [(gen << 24) | (nxt + i) for i in range(k)]
            nxt += k
            for i, h in enumerate(hs_):
                e.append(h)
                par.append(INVALID if i == 0 else hs_[i - 1 if k > 32 else int(rng.integers(0, i))])
                gid.append(g)
        if not e:
            return
        e, par, gid = np.array(e, np.uint32), np.array(par, np.uint32), np.array(gid)
        rc, slot = d.spawn(e, par, gid + gcount)
        assert rc == 0
        assert len(np.unique(slot)) == len(slot)
        for h, g in zip(e, gid):
            naive.add(int(h))
            groups.setdefault(int(g) + gcount, []).append(int(h))
        gcount += n_groups

    sizes_all = list(range(1, 33)) + [40, 70]   # 40, 70: chains longer than a window
    spawn(400, sizes_all, 0)
    for frame in range(60):
        keys = list(groups.keys())
        if frame % 3 == 0:   # a long run of neighbours in spawn order
            a = int(rng.integers(0, max(1, len(keys) - 30)))
            victims = keys[a:a + 30]
        else:
            victims = list(rng.choice(keys, max(1, len(keys) // 8), replace=False))
        dead = [h for g in victims for h in groups.pop(int(g))]
        d.despawn(np.array(dead, np.uint32))
        for h in dead:
            naive.remove(h)
        # refill to the brim every few frames, with whatever sizes
        sizes = sizes_all if frame % 2 else [1, 2, 3, 7, 32]
        spawn(10_000 if frame % 4 == 0 else len(victims), sizes, 1 + frame % 100)
        assert np.array_equal(d.dense_in_rank_order(), np.array(naive.dense, np.uint32)), f"frame {frame}"
        d.check_invariants()
        assert d.count <= cap and d.extent <= cap
    assert d.count > cap - 40   # the pool did get full
    d.close()
