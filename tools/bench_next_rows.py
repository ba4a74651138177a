#!/usr/bin/env python
"""Timings of the SURVEY.md 8(f) rows N1-N3 on one B200 (python tools/bench_next_rows.py > gpurun_out/next_rows.json).

Each row is timed through the C ABI from HOST buffers (wall clock around the call + a synchronise, median of 5), beside
the CPU work it replaces where the reference's code for it is compiled in oracle/_ref (N3: sc_world::ReadSectorFile)."""
import ctypes as C
import json
import statistics
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import scgpu  # noqa: E402
from scgpu import scenes  # noqa: E402


def med(fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e3


def main():
    out = {}
    n = 1_000_000
    vps = scenes.standard_views(1)
    # ---- N1: renderer sort + runs over 1 M draws (frozen culling: every instance is a draw) ------------------
    sc = scenes.city_flat(n)
    rng = np.random.default_rng(1)
    mm = np.stack([rng.integers(0, 64, n), rng.integers(0, 256, n)], axis=1).astype(np.uint32)
    s = scgpu.Scene(n, max_views=1, max_entity_index=n)
    e = np.arange(n, dtype=np.uint32)
    s.spawn(e, sc["trs9"], None, sc["aabb6"], mm, sc["flags"])
    s.set_views(vps)
    s.update(scgpu.UPDATE_FREEZE_CULLING)
    s.counts()
    mat_pipe = rng.integers(0, 2, 256).astype(np.uint32)
    kept, runs = C.c_uint32(0), C.c_uint32(0)

    def n1():
        s._ck(s.lib.scgpuBuildSortedDraws(s.ctx, 0, 0, mat_pipe.ctypes.data_as(C.c_void_p), 256, 64, None, C.byref(kept), None,
                                          C.byref(runs)), "sort")
    n1()
    out["N1_sorted_draws"] = {"draws": n, "runs": runs.value, "ms": med(n1),
                              "what": "scgpuBuildSortedDraws: keys + stable 8-bit counting-sort passes over the key bits in use (15 here: 2 passes) + runs + gather of 80-byte items, device-resident; hand-written, no library sort"}
    from oracle_bind import ref_available, ref_renderer_submit
    draws, _, _ = s.read_draw_items(0, 0)
    if ref_available():
        # the reference's own block of sc_vk.cpp (:1841-1912: filter, std::sort, bind-on-change loop) compiled into oracle/_ref
        t0 = time.perf_counter()
        ro, rb = ref_renderer_submit(draws, mat_pipe, 64)
        out["N1_sorted_draws"]["reference_loop_ms"] = (time.perf_counter() - t0) * 1e3
        out["N1_sorted_draws"]["reference_loop_bind_points"] = int((rb != 0).sum())
    t0 = time.perf_counter()
    order = np.lexsort((draws["meshId"], draws["materialId"], mat_pipe[draws["materialId"]]))
    out["N1_sorted_draws"]["numpy_lexsort_ms"] = (time.perf_counter() - t0) * 1e3
    s.close()

    # ---- N2: 38 k procedural sectors (~1 M instances) generated on the device ----------------------------------
    side = 196
    gx, gz = np.meshgrid(np.arange(side, dtype=np.int32) - side // 2, np.arange(side, dtype=np.int32) - side // 2, indexing="xy")
    coords = np.ascontiguousarray(np.stack([gx.ravel(), gz.ravel()], axis=1))
    gen = scgpu.SectorGen(sectorSizeMeters=64.0, seed=424242, propsPerSectorMin=18, propsPerSectorMax=34, includeGroundPlane=1,
                          meshCube=1, meshTriangle=2, matUnlit=1, matChecker=2, matTest=3)
    lib = scgpu.load_library()
    total = int(sum(lib.scgpuSectorSpawnCount(gen, int(x), int(z)) for x, z in coords))
    ent = np.arange(total, dtype=np.uint32)

    def n2():
        t = scgpu.Scene(total, max_views=1, max_entity_index=total)
        t0 = time.perf_counter()
        t.spawn_sectors(gen, coords, ent)
        t.synchronize()
        dt = time.perf_counter() - t0
        t.close()
        return dt
    n2()
    out["N2_spawn_sectors"] = {"sectors": int(coords.shape[0]), "instances": total, "ms": statistics.median([n2() for _ in range(5)]) * 1e3,
                               "h2d_bytes": int(coords.nbytes + ent.nbytes + 4 * (coords.shape[0] + 1)),
                               "vs_host_path_h2d_bytes": int(total * (4 + 36 + 8)),
                               "what": "scgpuSpawnSectors incl. host handle bookkeeping and H2D of coordinates + handles"}
    t0 = time.perf_counter()
    scenes.city_props(total)
    out["N2_spawn_sectors"]["numpy_generator_ms"] = (time.perf_counter() - t0) * 1e3

    # ---- N3: a 1 M-instance .scsector image unpacked on the device vs the reference's reader -------------------
    import oracle_bind
    if oracle_bind.ref_available():
        L = oracle_bind.ref_lib()
        L.screfWriteSectorFile.restype = C.c_int
        L.screfWriteSectorFile.argtypes = [C.c_char_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32]
        L.screfReadSectorInstances.restype = C.c_int
        L.screfReadSectorInstances.argtypes = [C.c_char_p, C.c_uint32] + [C.c_void_p] * 5
        f = lambda a: a.ctypes.data_as(C.c_void_p)
        trs = np.ascontiguousarray(sc["trs9"])
        ids = rng.integers(1, 1 << 40, n, dtype=np.uint64)
        mesh = rng.integers(1, 3, n).astype(np.uint64)
        mat = rng.integers(1, 4, n).astype(np.uint64)
        tags = np.zeros(n, np.uint32)
        with tempfile.TemporaryDirectory() as d:
            path = str(Path(d) / "big.scsector").encode()
            assert L.screfWriteSectorFile(path, 4, 0, 0, n, f(ids), f(ids), f(mesh), f(mat), f(trs), f(tags), 1)
            raw = Path(path.decode()).read_bytes()
            oxz = np.zeros(2, np.int32); oid = np.zeros(n, np.uint64); om = np.zeros(n, np.uint64); ot = np.zeros(n, np.uint64)
            otrs = np.zeros((n, 9), np.float32)
            ref_ms = med(lambda: L.screfReadSectorInstances(path, n, f(oxz), f(oid), f(om), f(ot), f(otrs)), 3)
        tab = scgpu.make_asset_table({1: 1, 2: 2}, 1, {1: 1, 2: 2, 3: 3}, 1)
        rawarr = np.frombuffer(raw, np.uint8)

        def n3():
            t = scgpu.Scene(n, max_views=1, max_entity_index=n)
            t0 = time.perf_counter()
            t._ck(t.lib.scgpuSpawnSectorFile(t.ctx, rawarr.ctypes.data_as(C.c_void_p), rawarr.size, e.ctypes.data_as(C.c_void_p), n,
                                             C.byref(tab)), "file")
            t.synchronize()
            dt = time.perf_counter() - t0
            t.close()
            return dt
        n3()
        out["N3_sector_file"] = {"instances": n, "file_bytes": len(raw), "ms": statistics.median([n3() for _ in range(5)]) * 1e3,
                                 "reference_ReadSectorFile_ms": ref_ms,
                                 "what": "scgpuSpawnSectorFile from a host file image (chunk walk + H2D of the INST payload + unpack kernel) vs "
                                         "sc_world::ReadSectorFile alone (the reference then still runs readSectorFile + World::add per record)"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
