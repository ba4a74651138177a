"""Host mirror of the Transform pool (sc-gameengine_b200/csrc/scgpu_pool.h, the host half of scgpuSpawn / scgpuDespawn)
compiled through tests/hostsim and checked on the CPU:
  * against a naive replay of ComponentPool::remove (sc_ecs.h:228-247) — dense order, sparse table, removed set;
  * the net moves, applied to a copy of the old dense array in ANY order, give the new dense array (sources lie in
    the vacated tail, never in a destination) — this is what k_despawn_apply relies on;
  * against the reference's own World (oracle/_ref) when it is built: same dense order after batches of destroys;
  * stale / unknown / repeated / invalid handles are skipped, a failed spawn leaves the pool untouched."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_bind as ob

HS_DIR = Path(__file__).resolve().parent / "hostsim"
INVALID = 0xFFFFFFFF


@pytest.fixture(scope="module")
def hs():
    subprocess.run(["make", "-C", str(HS_DIR)], check=True, capture_output=True)
    L = C.CDLL(str(HS_DIR / "libhostsim.so"))
    L.hs_pool_create.restype = C.c_void_p
    L.hs_pool_create.argtypes = [C.c_uint32]
    L.hs_pool_destroy.argtypes = [C.c_void_p]
    L.hs_pool_set_threads.argtypes = [C.c_void_p, C.c_uint32]
    L.hs_pool_spawn.restype = C.c_int
    L.hs_pool_spawn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.hs_pool_despawn.restype = C.c_double
    L.hs_pool_despawn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    for f in ("hs_pool_count", "hs_pool_num_moves", "hs_pool_num_removed"):
        getattr(L, f).restype = C.c_uint32
        getattr(L, f).argtypes = [C.c_void_p]
    L.hs_pool_read.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    return L


class Pool:
    def __init__(self, L, sparse_size, threads=1):
        self.L, self.sparse_size = L, sparse_size
        self.p = L.hs_pool_create(sparse_size)
        L.hs_pool_set_threads(self.p, threads)

    def close(self):
        self.L.hs_pool_destroy(self.p)

    def spawn(self, e):
        e = np.ascontiguousarray(e, np.uint32)
        bad = C.c_uint32(0)
        return self.L.hs_pool_spawn(self.p, len(e), e.ctypes.data, C.byref(bad)), bad.value

    def despawn(self, e):
        e = np.ascontiguousarray(e, np.uint32)
        sec = self.L.hs_pool_despawn(self.p, len(e), e.ctypes.data)
        moves = np.zeros((self.L.hs_pool_num_moves(self.p), 2), np.uint32)
        removed = np.zeros(self.L.hs_pool_num_removed(self.p), np.uint32)
        self.L.hs_pool_read(self.p, None, None, moves.ctypes.data, removed.ctypes.data)
        return moves, removed, sec

    def state(self):
        dense = np.zeros(self.L.hs_pool_count(self.p), np.uint32)
        sparse = np.zeros(self.sparse_size, np.uint32)
        self.L.hs_pool_read(self.p, dense.ctypes.data, sparse.ctypes.data, None, None)
        return dense, sparse


def naive_despawn(dense, victims):
    """ComponentPool::remove, one handle after the other (sc_ecs.h:228-247)."""
    dense = [int(x) for x in dense]
    where = {e: i for i, e in enumerate(dense)}
    removed = []
    for h in (int(v) for v in victims):
        s = where.get(h, -1)
        if h == INVALID or s < 0:
            continue
        last = len(dense) - 1
        if s != last:
            dense[s] = dense[last]
            where[dense[s]] = s
        dense.pop()
        del where[h]
        removed.append(h & 0xFFFFFF)
    return np.array(dense, np.uint32), np.array(removed, np.uint32)


def check_batch(pool, victims):
    before, _ = pool.state()
    moves, removed, _ = pool.despawn(victims)
    after, sparse = pool.state()
    want, want_removed = naive_despawn(before, victims)
    assert np.array_equal(after, want)
    assert np.array_equal(removed, want_removed)
    # sparse table == inverse of the dense array, nothing else set
    expect = np.zeros_like(sparse)
    expect[after & 0xFFFFFF] = np.arange(1, len(after) + 1, dtype=np.uint32)
    assert np.array_equal(sparse, expect)
    # the net moves reproduce the new order from the old one, in any application order
    if len(moves):
        assert len(np.unique(moves[:, 0])) == len(moves) == len(np.unique(moves[:, 1]))
        assert moves[:, 0].max() < len(after) <= moves[:, 1].min()
        assert moves[:, 1].max() < len(before)
    replay = before.copy()
    for d, s in moves[::-1]:
        replay[d] = before[s]
    assert np.array_equal(replay[: len(after)], after)


def handles(idx, gen=0):
    return (np.asarray(idx, np.uint32) & np.uint32(0xFFFFFF)) | np.uint32(gen << 24)


def test_single_and_edge_batches(hs):
    p = Pool(hs, 64)
    assert p.spawn(handles(range(40))) == (0, 0)
    check_batch(p, handles([39]))                 # the last element: no move
    check_batch(p, handles([0]))                  # the first: tail moves in
    check_batch(p, handles([5, 5, 5]))            # repeated handle: second and third are stale
    check_batch(p, handles([63, 50]))             # never spawned
    check_batch(p, np.array([INVALID, 0x01000003, 7 | (200 << 24)], np.uint32))  # invalid handle, wrong generations
    check_batch(p, np.array([0x00FFFFFF], np.uint32))  # index beyond the sparse table
    check_batch(p, np.zeros(0, np.uint32))
    d, _ = p.state()
    check_batch(p, d[::-1].copy())                # everything, back to front
    assert hs.hs_pool_count(p.p) == 0
    assert p.spawn(handles(range(10), gen=1)) == (0, 0)
    d, _ = p.state()
    check_batch(p, d.copy())                      # everything, front to back (every removal moves the tail)
    p.close()


def test_tail_chains(hs):
    """victims that sit in the tail get swapped into holes before their own turn comes: chains of moves"""
    p = Pool(hs, 256)
    p.spawn(handles(range(200)))
    check_batch(p, handles([0, 199, 1, 198, 2, 197, 100, 196, 195, 3]))
    check_batch(p, handles(list(range(150, 190)) + list(range(10, 50))))
    p.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_churn_against_naive(hs, seed):
    rng = np.random.default_rng(seed)
    cap = 4096
    p = Pool(hs, cap)
    gen = np.zeros(cap, np.uint32)
    free = list(range(cap))
    rng.shuffle(free)
    for it in range(12):
        dense, _ = p.state()
        k = min(len(free), int(rng.integers(100, 900)))
        idx = np.array([free.pop() for _ in range(k)], np.uint32)
        assert p.spawn(idx | (gen[idx] << np.uint32(24))) == (0, 0)
        dense, _ = p.state()
        m = int(rng.integers(1, max(2, len(dense) // 2)))
        victims = rng.choice(dense, m, replace=False)
        if it % 3 == 0:   # grouped victims (whole runs of slots) like sector unloads
            a = int(rng.integers(0, len(dense) - 1))
            victims = dense[a: a + m].copy()
        noise = rng.integers(0, 1 << 32, 17, dtype=np.uint64).astype(np.uint32)
        batch = np.concatenate([victims, noise, victims[:5]])
        rng.shuffle(batch)
        before = set(int(x) for x in dense)
        check_batch(p, batch)
        after, _ = p.state()
        for h in before - set(int(x) for x in after):
            i = h & 0xFFFFFF
            gen[i] = (gen[i] + 1) & 0xFF
            free.append(i)
    p.close()


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_large_batches_threaded_gather_and_scatter(hs, threads):
    """batches above the threading threshold (32 Ki handles): the result must not depend on the thread count"""
    rng = np.random.default_rng(7)
    n = 200_000
    p = Pool(hs, 1 << 18, threads)
    assert p.spawn(handles(rng.permutation(n))) == (0, 0)
    for it in range(4):
        dense, _ = p.state()
        victims = rng.choice(dense, 50_000 - 7000 * it, replace=False)
        if it == 1:   # mostly tail elements, in pool order: long move chains
            victims = dense[-45_000:].copy()
        # frame 0 names no element twice (victims compacted in parallel), the others do (sequential walk, first wins)
        batch = np.concatenate([victims, victims[:3000] if it else victims[:0],
                                rng.integers(1 << 30, 1 << 32, 500, dtype=np.uint64).astype(np.uint32)])
        if it >= 2:
            rng.shuffle(batch)
        check_batch(p, batch)
    p.close()


def test_failed_spawn_leaves_pool_untouched(hs):
    p = Pool(hs, 32)
    assert p.spawn(handles([1, 2, 3])) == (0, 0)
    d0, s0 = p.state()
    assert p.spawn(handles([4, 5, 2, 6])) == (3, 2)          # index 2 already owns a Transform
    assert p.spawn(np.array([7, INVALID], np.uint32)) == (1, 1)
    assert p.spawn(handles([8, 40])) == (2, 1)               # beyond max_entity_index
    assert p.spawn(handles([9, 9])) == (3, 1)                # duplicate inside the batch
    d1, s1 = p.state()
    assert np.array_equal(d0, d1) and np.array_equal(s0, s1)
    assert p.spawn(handles([4, 5, 6])) == (0, 0)
    assert np.array_equal(p.state()[0], handles([1, 2, 3, 4, 5, 6]))
    p.close()


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built")
def test_same_order_as_the_reference_pool(hs):
    """The reference's own World: create + add<Transform>, batches of destroy, compare m_denseEntities."""
    rng = np.random.default_rng(11)
    ref = ob.RefScene()
    p = Pool(hs, 1 << 16)
    trs = np.tile(np.array([0, 0, 0, 0, 0, 0, 1, 1, 1], np.float32), (4000, 1))
    for it in range(6):
        e = ref.create_entities(3000 + 100 * it)
        ref.spawn(e, trs[: len(e)])
        assert p.spawn(e) == (0, 0)
        dense = ref.dense_entities()
        assert np.array_equal(dense, p.state()[0])
        victims = rng.choice(dense, len(dense) // 3, replace=False)
        if it % 2:
            victims = np.sort(victims)
        ref.despawn(victims)
        p.despawn(np.concatenate([victims, victims[:7]]))    # the repeats are stale by then, as in World::destroy
        assert np.array_equal(ref.dense_entities(), p.state()[0])
    ref.close()
    p.close()
