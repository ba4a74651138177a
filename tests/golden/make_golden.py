#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libscref.so = the unmodified sources under
/root/reference compiled headless, see oracle/Makefile). Run in the build container only:

    GLIBC_TUNABLES=glibc.cpu.hwcaps=-FMA,-AVX2 python tests/golden/make_golden.py

The tunable pins glibc's generic (non-FMA) sinf/cosf variant, the libm the oracle restates.
The fixtures are derived data (inputs + the reference's outputs), never reference code. They travel to the GPU box,
where /root/reference does not exist.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))

import oracle_bind  # noqa: E402
from oracle_bind import RefScene  # noqa: E402
from scenarios import INVALID, random_aabb, random_forest, random_trs  # noqa: E402
from scgpu import scenes  # noqa: E402


def f(a):
    return a.ctypes.data_as(C.c_void_p)


def kats(L):
    out = {}
    m = np.zeros(16, np.float32)
    p, r, s = (np.array(x, np.float32) for x in ((1, 2, 3), (.1, .2, .3), (2, 3, 4)))
    L.screfMat4Trs(f(p), f(r), f(s), f(m))
    out["trs_child"] = m.copy()
    pm = np.zeros(16, np.float32)
    p2, r2, s2 = (np.array(x, np.float32) for x in ((10, 0, -5), (0, 1.5, 0), (1, 1, 1)))
    L.screfMat4Trs(f(p2), f(r2), f(s2), f(pm))
    out["trs_parent"] = pm.copy()
    w = np.zeros(16, np.float32)
    L.screfMat4Mul(f(pm), f(m), f(w))
    out["parent_x_child"] = w.copy()
    # random TRS / mul / inverse / frustum / sphere / sphere-in-frustum vectors
    rng = np.random.default_rng(20261018)
    n = 512
    trs = random_trs(rng, n, spread=100.0)
    trs[0, 3:6] = 0
    trs[1, 3:6] = [1e-5, -1e-5, 0.7853981]
    trs[2, 3:6] = [119.99, 120.0, -1e6]
    trs[3, 3:6] = [3.14159265, -3.14159265, 6.2831853]
    trs_out = np.zeros((n, 16), np.float32)
    for i in range(n):
        L.screfMat4Trs(f(trs[i, 0:3].copy()), f(trs[i, 3:6].copy()), f(trs[i, 6:9].copy()), f(trs_out[i]))
    out["rand_trs_in"], out["rand_trs_out"] = trs, trs_out
    a = rng.normal(size=(n, 16)).astype(np.float32)
    b = rng.normal(size=(n, 16)).astype(np.float32)
    mul = np.zeros((n, 16), np.float32)
    inv = np.zeros((n, 16), np.float32)
    pl = np.zeros((n, 24), np.float32)
    bb = np.sort(rng.normal(size=(n, 2, 3)).astype(np.float32), axis=1).reshape(n, 6)
    sph = np.zeros((n, 4), np.float32)
    inside = np.zeros(n, np.int32)
    for i in range(n):
        L.screfMat4Mul(f(a[i]), f(b[i]), f(mul[i]))
        L.screfMat4Inverse(f(a[i]), f(inv[i]))
        L.screfFrustumFromViewProj(f(a[i]), f(pl[i]))
        L.screfWorldBoundsSphere(f(a[i]), f(bb[i]), f(sph[i]), f(sph[i, 3:]))
        inside[i] = L.screfSphereInFrustum(f(pl[i]), f(sph[i, :3].copy()), C.c_float(float(sph[i, 3])))
    out.update(mat_a=a, mat_b=b, mul=mul, inverse=inv, planes=pl, aabb=bb, sphere=sph, inside=inside)
    # sinf / cosf: strided sweeps over all float bit patterns, hashed; plus 4096 raw samples
    hs, hc = C.c_uint64(), C.c_uint64()
    L.screfSinCosSweep(0, (1 << 32) // 4099 + 1, 4099, C.byref(hs), C.byref(hc))
    out["sweep_stride"] = np.array([4099], np.uint64)
    out["sweep_hash"] = np.array([hs.value, hc.value], np.uint64)
    bits = (np.arange(4096, dtype=np.uint64) * 1048583 + 12345).astype(np.uint32)
    xs = bits.view(np.float32)
    out["sc_x"] = xs
    out["sc_sin"] = np.array([L.screfSinf(C.c_float(float(x))) for x in xs], np.float32)
    out["sc_cos"] = np.array([L.screfCosf(C.c_float(float(x))) for x in xs], np.float32)
    return out


def snapshot(r, entities, n_views_vp, max_draws=(0,)):
    """everything the parity tests compare, as produced by the reference"""
    e = np.ascontiguousarray(entities, np.uint32)
    n = len(e)
    par = np.zeros(n, np.uint32)
    trs = np.zeros((n, 9), np.float32)
    dirty = np.zeros(n, np.uint8)
    r.L.screfReadTransform(r.w, n, f(e), f(par), f(trs), f(dirty))
    fl = np.zeros(n, np.uint32)
    bb = np.zeros((n, 6), np.float32)
    mm = np.zeros((n, 2), np.uint32)
    r.L.screfReadComponents(r.w, n, f(e), f(fl), f(bb), f(mm))
    d = dict(entity=e, parent_after=par, trs_after=trs, flags=fl, aabb=bb, mesh_mat=mm, world=r.read_world(e),
             view_proj=np.ascontiguousarray(n_views_vp, np.float32))
    for v, lst in enumerate(r.visible):
        d[f"visible_{v}"] = lst
        d[f"culled_{v}"] = r.culled[v]
    for md in max_draws:
        items, em, dr = r.read_draw_items(0, md)
        d[f"draws_{md}_entity"] = items["entity"].copy()
        d[f"draws_{md}_mesh"] = items["meshId"].copy()
        d[f"draws_{md}_mat"] = items["materialId"].copy()
        d[f"draws_{md}_model"] = items["model"].copy()
        d[f"draws_{md}_stats"] = np.array([em, dr], np.uint32)
    return d


def default_scene(L):
    """Config 1: the sandbox's default streamed scene (src/sandbox/src/main.cpp:66-99) after 60 frames."""
    r = RefScene(0)
    sectors = L.screfBuildDefaultScene(r.w, 60)
    e = r.dense_entities()
    vp = np.zeros(16, np.float32)
    L.screfGetViewProj(r.w, f(vp))
    n = len(e)
    vis = np.zeros(n, np.uint32)
    cul = np.zeros(n, np.uint32)
    nv = L.screfReadVisible(r.w, n, f(vis))
    nc = L.screfReadCulled(r.w, n, f(cul))
    r.visible, r.culled = [vis[:nv].copy()], [cul[:nc].copy()]
    d = snapshot(r, e, vp.reshape(1, 16), max_draws=(0, 100))
    d["planes"] = r.planes()
    d["sectors"] = np.array([sectors], np.uint32)
    # inputs as they were BEFORE the last TransformSystem ran are not recoverable; the scene is static, so the
    # post-state (trs_after/parent_after, all clean) with every instance re-marked dirty reproduces the same result.
    r.close()
    return d


def forest_scene(seed):
    """seeded random forest with the edge cases of SURVEY.md §3.2/§3.4, three frames (spawn / edit / churn)"""
    rng = np.random.default_rng(seed)
    r = RefScene(0)
    n = 1500
    e = r.create_entities(n)
    parent_idx = random_forest(rng, n, max_back=600)
    trs = random_trs(rng, n, spread=60.0)
    flags = rng.choice([0, 1, 2, 3], size=n, p=[0.05, 0.1, 0.15, 0.7]).astype(np.uint32)
    trs[5, 6:9] = 0.0
    trs[6, 6:8] = 0.0
    trs[7, 0] = np.nan
    trs[8, 4] = np.inf
    trs[9, 3:6] = [1e6, -3e9, 1e-30]
    trs[10, 3:6] = [120.0, -119.99, 0.78539819]
    trs[11, 6] = -2.0
    par = scenes.parent_handles(parent_idx, e)
    par[20] = e[20]
    par[21] = 0x00ABCDEF
    par[30], par[31] = e[31], e[30]
    par[40], par[41], par[42] = e[41], e[42], e[40]
    par[43] = e[40]
    par[300], par[900] = e[900], e[300]
    par[901] = e[300]
    aabb = random_aabb(rng, n)
    mm = rng.integers(0, 50, size=(n, 2)).astype(np.uint32)
    vps = scenes.standard_views(3, center=(0.0, 10.0, 80.0))
    out = dict(in_entity=e, in_trs=trs, in_parent=par, in_aabb=aabb, in_mesh_mat=mm, in_flags=flags, view_proj=vps)
    r.spawn(e, trs, par, aabb, mm, flags)
    r.update(vps)
    for k, v in snapshot(r, e, vps, (0, 13)).items():
        out["f0_" + k] = v
    idx = rng.choice(n, 200, replace=False)
    t2 = random_trs(rng, 200, spread=60.0)
    sp_e = e[[40, 50, 51]]
    sp_p = np.array([INVALID, e[52], e[20]], np.uint32)
    r.set_local(e[idx], t2)
    r.set_parent(sp_e, sp_p)
    r.update(vps)
    out.update(f1_set_entity=e[idx], f1_set_trs=t2, f1_setparent_entity=sp_e, f1_setparent_parent=sp_p)
    for k, v in snapshot(r, e, vps, (0,)).items():
        out["f1_" + k] = v
    dead = rng.choice(n, 300, replace=False)
    dead_handles = np.concatenate([e[dead], np.array([0x00FFFFF0, e[dead[0]]], np.uint32)])
    r.despawn(dead_handles)
    e2 = r.create_entities(100)
    live = np.setdiff1d(np.arange(n), dead)
    trs3 = random_trs(rng, 100, spread=60.0)
    par3 = e[rng.choice(live, 100)]
    bb3 = random_aabb(rng, 100)
    r.spawn(e2, trs3, par3, bb3, None, None)
    r.update(vps)
    out.update(f2_despawn=dead_handles, f2_spawn_entity=e2, f2_spawn_trs=trs3, f2_spawn_parent=par3, f2_spawn_aabb=bb3,
               f2_dense=r.dense_entities())
    alive = np.concatenate([e[live], e2])
    for k, v in snapshot(r, alive, vps, (0, 13)).items():
        out["f2_" + k] = v
    r.close()
    return out


def traffic_golden(L):
    """the reference's TrafficAISystem over its own procedural lanes + a junction-rich random graph"""
    from oracle_bind import LANE_KEYS, RefLanes, ref_traffic_frames
    g0 = scenes.lane_random(120, 320, seed=77, hostile=False)
    lanes = RefLanes(3.5, 12.0)
    for i in range(len(g0["node_speed"])):
        d = np.array([(i % 1000 - 500) / 1000.0, 0.1, 0.25], np.float32)
        assert lanes.add_node(g0["node_pos"][i], d, float(g0["node_speed"][i])) == i
    for s in range(len(g0["seg_len"])):
        lanes.add_segment(int(g0["seg_nodes"][s, 0]), int(g0["seg_nodes"][s, 1]), g0["seg_dir"][s])
    for sx in range(3):
        for sz in range(3):
            lanes.build_sector(sx + 10, sz + 10, 64.0)
    lanes.remove_sector(11, 11)
    for s in (5, 17, 100, 211):
        lanes.set_active(s, False)
    g = lanes.export()
    agents, trs = scenes.traffic_agents(g, 400, seed=9)
    dts = np.array([1 / 60] * 12 + [0.0, 0.5, 2.5, 1 / 144, 1 / 30, 1 / 60], np.float32)
    frames = ref_traffic_frames(lanes, agents, trs, [float(x) for x in dts])
    lanes.close()
    out = {"g_" + k: g[k] for k in LANE_KEYS}
    out["g_default_speed"] = g["default_speed"]
    out.update(a_lane=agents["lane"], a_s=agents["s"], a_speed=agents["speed"], a_look=agents["look"], trs=trs, dts=dts)
    for i, k in enumerate(("lane", "s", "speed", "look", "trs", "dirty")):
        out["f_" + k] = np.stack([fr[i] for fr in frames])
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.normal(size=300) * 10.0 ** rng.integers(-6, 3, 300), [0, -0.0, 88, 89, -104, -87.4, 1e-30, np.inf, -np.inf, np.nan,
                        0.4375, 0.6875, 1.1875, 2.4375, 2.0 ** 25, 1.0, -1.0]]).astype(np.float32)
    out["kat_x"] = x
    out["kat_expf"] = np.array([L.screfExpf(float(v)) for v in x], np.float32)
    out["kat_atanf"] = np.array([L.screfAtanf(float(v)) for v in x], np.float32)
    y2 = np.concatenate([rng.normal(size=300), [0, -0.0, 1, -1, np.inf, -np.inf, 0, 1e-40, 1e30, 3]]).astype(np.float32)
    x2 = np.concatenate([rng.normal(size=300), [-1, -2, 0, -0.0, np.inf, -np.inf, 0, 1e30, 1e-40, 1]]).astype(np.float32)
    out["kat_y2"], out["kat_x2"] = y2, x2
    out["kat_atan2f"] = np.array([L.screfAtan2f(float(a), float(b)) for a, b in zip(y2, x2)], np.float32)
    return out


def main():
    if "hwcaps=-FMA" not in os.environ.get("GLIBC_TUNABLES", ""):
        env = dict(os.environ, GLIBC_TUNABLES="glibc.cpu.hwcaps=-FMA,-AVX2")
        os.execve(sys.executable, [sys.executable] + sys.argv, env)
    assert oracle_bind.ref_available(), "build oracle/_ref first: make -C oracle ref"
    L = oracle_bind.ref_lib()
    if sys.argv[1:] == ["traffic"]:  # only the traffic fixture (the others are unchanged)
        np.savez_compressed(HERE / "traffic.npz", **traffic_golden(L))
        return
    np.savez_compressed(HERE / "traffic.npz", **traffic_golden(L))
    np.savez_compressed(HERE / "kats.npz", **kats(L))
    np.savez_compressed(HERE / "default_scene.npz", **default_scene(L))
    np.savez_compressed(HERE / "forest_scene.npz", **forest_scene(4242))
    for p in sorted(HERE.glob("*.npz")):
        print(p.name, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
