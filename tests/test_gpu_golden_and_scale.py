"""GPU path against the committed golden fixtures (outputs of the reference itself), plus size-independent
properties at BASELINE.json's full sizes, plus the NCCL gather across GPUs when more than one is visible."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from scenarios import (GpuAdapter, assert_same_bits, check_snapshot, load_golden, replay_default_scene,
                       replay_forest_scene)
from scgpu import scenes

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_default_sandbox_scene_golden_on_gpu():
    """Config 1: 721 transforms / 719 renderables -> 166 visible, 553 culled, 166 draws, bit for bit"""
    g = GpuAdapter(1024, max_views=1)
    replay_default_scene(g, load_golden("default_scene.npz"))
    c = g.s.counts()
    assert (c.transforms, c.renderablesTotal, c.visible[0], c.culled[0]) == (721, 719, 166, 553)
    assert np.array_equal(g.s.get_view_planes(0).view(np.uint32), load_golden("default_scene.npz")["planes"].view(np.uint32))
    g.close()


def test_forest_scene_golden_on_gpu():
    gold = load_golden("forest_scene.npz")
    g = GpuAdapter(4096, max_views=3)
    replay_forest_scene(g, gold)
    assert np.array_equal(g.dense_entities(), gold["f2_dense"])
    assert np.array_equal(g.s.read_parents(gold["f2_entity"]), gold["f2_parent_after"])
    g.close()


def test_trs_golden_vectors_on_gpu():
    """512 reference mat4_trs results reproduced through the kernel (flat instances, world == local)"""
    k = load_golden("kats.npz")
    n = len(k["rand_trs_in"])
    g = GpuAdapter(n, max_views=1)
    e = np.arange(n, dtype=np.uint32)
    g.spawn(e, k["rand_trs_in"])
    g.update(np.eye(4, dtype=np.float32).ravel())
    assert_same_bits(g.read_world(e), k["rand_trs_out"], "mat4_trs golden")
    g.close()


@pytest.mark.parametrize("kind,n,views", [("flat", 1_000_000, 1), ("hier", 16 * 1024 * 1024 - 4096, 5)])
def test_full_size_properties(kind, n, views):
    """BASELINE.json configs[1] and configs[2] at full size: properties that need no CPU run.
       * idempotence: recomputing everything reproduces the same matrices and lists
       * freeze culling: every candidate visible, in pool order
       * counts add up; visible lists are strictly increasing in slot order (stable compaction)
       * a 200k-instance prefix cross-checked bit for bit against the oracle"""
    import scgpu
    from oracle_bind import PortScene
    sc = scenes.city_flat(n) if kind == "flat" else scenes.city_hier(n)
    e = np.arange(n, dtype=np.uint32)
    par = scenes.parent_handles(sc["parent"], e)
    vps = scenes.standard_views(views)
    s = scgpu.Scene(n, max_views=views, max_entity_index=n)
    s.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
    s.set_views(vps)
    s.update()
    c1 = s.counts()
    assert c1.transforms == n and c1.recomputed == n
    cand = int((sc["flags"] & 2 != 0).sum())
    assert c1.renderablesTotal == cand
    lists1 = [s.read_visible(v) for v in range(views)]
    for v in range(views):
        assert len(lists1[v]) == c1.visible[v] and c1.visible[v] + c1.culled[v] == cand
        assert np.all(np.diff(lists1[v].astype(np.int64)) > 0)  # entity == slot here: pool order, no duplicates
    sample = np.random.default_rng(1).choice(n, 50_000, replace=False).astype(np.uint32)
    w1 = s.read_world(sample)
    s.mark_all_dirty()
    s.update()
    assert s.counts().recomputed == n
    for v in range(views):
        assert np.array_equal(s.read_visible(v), lists1[v])
    assert_same_bits(s.read_world(sample), w1, "idempotent recompute")
    s.update(scgpu.UPDATE_FREEZE_CULLING)
    assert s.counts().recomputed == 0
    for v in range(views):
        assert np.array_equal(s.read_visible(v), e[sc["flags"] & 2 != 0])
    # prefix parity: groups never reference later slots' parents beyond the prefix cut by construction of the cut
    m = 200_000
    while m < n and sc["parent"][m] >= 0:
        m += 1
    p = PortScene()
    p.spawn(e[:m], sc["trs9"][:m], par[:m], sc["aabb6"][:m], sc["mesh_mat"][:m], sc["flags"][:m])
    p.update(vps)
    assert_same_bits(s.read_world(e[:m]), p.world, "prefix world matrices vs oracle")
    s.update()
    for v in range(views):
        lst = s.read_visible(v)
        assert np.array_equal(lst[lst < m], p.visible[v]), f"prefix visible list, view {v}"
    s.close()


def test_nccl_gather_across_gpus():
    """Real NCCL path: one process per GPU (torchrun), shard union == unsharded. Needs >= 2 GPUs on the box."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("single-GPU box")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(ROOT / "tests" / "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU OK" in r.stdout and "MULTIGPU PEER OK" in r.stdout and "MULTIGPU PEER RECOVERY OK" in r.stdout


def test_dropin_adapter_systems_against_reference_systems():
    """The C++ adapter systems (sc-gameengine_b200/host) and the reference's own CPU systems run on ONE sc::World
    (the sandbox's default streamed scene + scripted edits, churn, freeze, draw budget) and must produce the same
    CullingState, RenderFrameData and world matrices every frame. The binary holds the reference's compiled code, so
    it is prebuilt in the build container (oracle/Makefile `dropin`) and travels in oracle/_ref/."""
    exe = ROOT / "oracle" / "_ref" / "sc_dropin_test"
    if not exe.exists():
        pytest.skip("oracle/_ref/sc_dropin_test not built (needs /root/reference)")
    env = dict(os.environ, GLIBC_TUNABLES="glibc.cpu.hwcaps=-FMA,-AVX2")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300, env=env)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "DROPIN OK" in r.stdout


def test_dropin_adapter_at_scale_outputs_equal_and_host_cost():
    """sc_dropin_test --bench: the three adapter systems and the reference's three on one World of 200 k entities (groups of
    five), 10 % dirtied per frame through sc::setLocal: ordered visible lists, candidates and draw counts equal every
    frame, no resync, and the adapter's per-frame host cost (one pool walk against a dense shadow table, no hashing)
    below the reference's own TransformSystem."""
    import json
    exe = ROOT / "oracle" / "_ref" / "sc_dropin_test"
    if not exe.exists():
        pytest.skip("oracle/_ref/sc_dropin_test not built (needs /root/reference)")
    env = dict(os.environ, GLIBC_TUNABLES="glibc.cpu.hwcaps=-FMA,-AVX2")
    r = subprocess.run([str(exe), "--bench", "200000", "6"], capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(line[-1])
    print(res)
    assert res["outputs_equal"] and res["resyncs"] == 0
    assert res["adapter_ms"]["TransformSystem"] < res["reference_ms"]["TransformSystem"]


def _default_scene_sectors(g):
    """(coords in activation order, rows of each sector) of the reference's default scene: every streamed sector
    starts with its ground plane (scale 64 x 0.1 x 64 at the sector centre)."""
    t = g["trs_after"]
    ground = np.nonzero((t[:, 6] == 64.0) & (t[:, 7] == np.float32(0.10)) & (t[:, 8] == 64.0))[0]
    coords = np.stack([np.floor(t[ground, 0] / 64.0), np.floor(t[ground, 2] / 64.0)], axis=1).astype(np.int32)
    ends = np.append(ground[1:], len(t))
    return coords, ground, ends


def _default_sector_gen(g):
    import scgpu
    mm = g["mesh_mat"]
    coords, ground, ends = _default_scene_sectors(g)
    cube = int(mm[ground[0], 0])
    others = np.setdiff1d(np.unique(mm[ground[0]:, 0]), [cube])
    tri = int(others[0]) if len(others) else cube
    # WorldPartitionConfig of the sandbox (src/sandbox/src/main.cpp:74-99): 64 m, seed 424242, 18..34 props, ground plane
    return scgpu.SectorGen(sectorSizeMeters=64.0, seed=424242, propsPerSectorMin=18, propsPerSectorMax=34,
                           includeGroundPlane=1, meshCube=cube, meshTriangle=tri, matUnlit=0, matChecker=0, matTest=0)


def test_sector_spawn_on_device_reproduces_the_reference_scene():
    """the 25 sectors of the sandbox's default scene generated ON THE DEVICE from (seed, coordinate)
    — scgpuSpawnSectors — give the frame the reference produced from its own generateSectorSpawnsStatic + World::add
    loop: world matrices of all 721 entities, the ordered visible list and the draw items (golden fixture)."""
    import scgpu
    g = load_golden("default_scene.npz")
    coords, ground, ends = _default_scene_sectors(g)
    first = int(ground[0])
    e = g["entity"]
    gen = _default_sector_gen(g)
    a = GpuAdapter(1024, max_views=1)
    # camera, Root, TriangleEntity, CubeEntity (sc_ecs.cpp:320-365) come from the host as before
    a.spawn(e[:first], g["trs_after"][:first], g["parent_after"][:first], g["aabb"][:first], g["mesh_mat"][:first], g["flags"][:first])
    counts = [a.s.lib.scgpuSectorSpawnCount(gen, int(x), int(z)) for x, z in coords]
    assert counts == [int(b - s0) for s0, b in zip(ground, ends)]
    a.s.spawn_sectors(gen, coords, e[first:])
    a.update(g["view_proj"])
    check_snapshot(a, g, "", 1, "device-generated default scene")
    assert np.array_equal(a.dense_entities(), e)
    a.close()


def test_sector_spawn_on_device_equals_host_generated_city():
    """N2 with distinct mesh / material handles and no ground planes: the device generator against the numpy
    restatement of generateSectorSpawnsStatic that feeds every other test (scgpu.scenes.city_props)."""
    import scgpu
    n = 20_000
    p = scenes.city_props(n, seed=777)
    sec = p["sector"]
    change = np.nonzero(np.any(np.diff(sec, axis=0) != 0, axis=1))[0] + 1
    coords = sec[np.concatenate([[0], change])]
    gen = scgpu.SectorGen(sectorSizeMeters=64.0, seed=777, propsPerSectorMin=18, propsPerSectorMax=34, includeGroundPlane=0,
                          meshCube=1, meshTriangle=2, matChecker=1, matTest=2, matUnlit=3)
    dev = GpuAdapter(n + 64, max_views=1)
    counts = [dev.s.lib.scgpuSectorSpawnCount(gen, int(x), int(z)) for x, z in coords]
    total = int(np.sum(counts))
    assert total >= n
    e = np.arange(total, dtype=np.uint32)
    dev.s.spawn_sectors(gen, coords, e)
    host = GpuAdapter(n + 64, max_views=1)
    host.spawn(e[:n], p["trs9"], None, None, p["mesh_mat"], None)
    vps = scenes.standard_views(1)
    for s in (dev, host):
        s.update(vps, freeze=True)
    assert_same_bits(dev.read_world(e[:n]), host.read_world(e[:n]), "device-generated sectors")
    dd, _, _ = dev.read_draw_items(0, 0)
    hd, _, _ = host.read_draw_items(0, 0)
    assert np.array_equal(dd["meshId"][:n], hd["meshId"]) and np.array_equal(dd["materialId"][:n], hd["materialId"])
    # a wrong handle count is refused and leaves the scene untouched
    with pytest.raises(scgpu.ScGpuError):
        dev.s.spawn_sectors(gen, coords[:2], np.arange(total, total + 3, dtype=np.uint32))
    assert dev.s.counts().transforms == total
    dev.close()
    host.close()


def test_sector_files_unpacked_on_device_match_the_reference_reader():
    """.scsector images written by the reference (format versions 1, 3, 4; with lane / spawner chunks
    around the INST chunk) spawned through scgpuSpawnSectorFile give the frame that the instances READ BY THE
    REFERENCE'S OWN READER give when spawned the ordinary way: world matrices, lists, mesh / material handles."""
    import scgpu
    from oracle_bind import PortScene
    g = load_golden("sector_files.npz")
    ids = {str(n): int(i) for n, i in zip(g["asset_names"], g["asset_ids"])}
    mesh_handles = {ids["meshes/cube"]: 3, ids["meshes/triangle"]: 4}
    mat_handles = {ids["materials/unlit"]: 5, ids["materials/checker"]: 6, ids["materials/test"]: 7}
    tab = scgpu.make_asset_table(mesh_handles, 3, mat_handles, 5)
    dev, p = GpuAdapter(1024, max_views=1), PortScene()
    base = 0
    for k in range(int(g["n_files"])):
        n = len(g[f"f{k}_id"])
        e = np.arange(base, base + n, dtype=np.uint32)
        base += n
        dev.s.spawn_sector_file(g[f"f{k}_bytes"].tobytes(), e, tab)
        mm = np.stack([[0 if i == 0 else mesh_handles.get(int(i), 3) for i in g[f"f{k}_mesh"]],
                       [0 if i == 0 else mat_handles.get(int(i), 5) for i in g[f"f{k}_mat"]]], axis=1).astype(np.uint32).reshape(n, 2)
        if n:
            p.spawn(e, g[f"f{k}_trs"], None, None, mm, None)
    vps = scenes.standard_views(1, center=(0.0, 20.0, 0.0))
    for s in (dev, p):
        s.update(vps)
    allE = np.arange(base, dtype=np.uint32)
    from scenarios import compare_frame, compare_draws
    compare_frame(dev, p, allE, 1, "sector files")
    for s in (dev, p):
        s.update(vps, freeze=True)
    compare_draws(dev, p, 0, "sector files, every draw")
    with pytest.raises(scgpu.ScGpuError):   # handle count must match the file
        dev.s.spawn_sector_file(g["f1_bytes"].tobytes(), np.arange(base, base + 3, dtype=np.uint32), tab)
    dev.close()


def test_editor_draw_items_match_reference_trs_goldens():
    """SURVEY 8(f) N4 (the second mat4_trs caller): scgpuBuildEditorDraws against BuildDrawItems
    (editor_core.cpp:242-264) restated over the REFERENCE's own mat4_trs outputs (tests/golden/kats.npz): entities
    without a mesh or a material handle are skipped, document order is kept, model matrices are bit-identical,
    flags are 0; plus hostile values (NaN / Inf / huge scale) against the plain-C oracle."""
    import ctypes as C
    import oracle_bind
    import scgpu
    k = load_golden("kats.npz")
    trs, want = k["rand_trs_in"], k["rand_trs_out"]
    n = len(trs)
    rng = np.random.default_rng(3)
    mesh = rng.integers(0, 4, n).astype(np.uint64) * np.uint64(0x100000001)      # 0 = no mesh; 64-bit handles
    mat = rng.integers(0, 3, n).astype(np.uint64) * np.uint64(0x200000003)
    s = scgpu.Scene(16, max_views=1)
    got = s.editor_draws(trs, mesh, mat)
    keep = (mesh != 0) & (mat != 0)
    assert 0 < keep.sum() < n and len(got) == keep.sum()
    assert np.array_equal(got["mesh"], mesh[keep]) and np.array_equal(got["material"], mat[keep])
    assert np.all(got["flags"] == 0)
    assert_same_bits(got["model"], want[keep], "editor draw models")
    # hostile inputs take the dense path: compare with the oracle's mat4_trs
    port = oracle_bind.port_lib()
    bad = np.tile(np.array([1, 2, 3, .1, .2, .3, 1, 1, 1], np.float32), (6, 1))
    bad[0, 0] = np.nan; bad[1, 4] = np.inf; bad[2, 6] = 3e38; bad[3, 6:9] = 0; bad[4, 3] = 1e30; bad[5, 7] = -0.0
    exp = np.zeros((6, 16), np.float32)
    f = lambda a: a.ctypes.data_as(C.c_void_p)
    for i in range(6):
        port.sco_mat4_trs(f(bad[i, 0:3].copy()), f(bad[i, 3:6].copy()), f(bad[i, 6:9].copy()), f(exp[i]))
    got2 = s.editor_draws(bad, np.ones(6, np.uint64), np.ones(6, np.uint64))
    assert_same_bits(got2["model"], exp, "editor draw models, hostile values")
    assert len(s.editor_draws(np.zeros((0, 9), np.float32), np.zeros(0, np.uint64), np.zeros(0, np.uint64))) == 0
    s.close()
