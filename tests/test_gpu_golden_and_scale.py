"""GPU path against the committed golden fixtures (outputs of the reference itself), plus size-independent
properties at BASELINE.json's full sizes, plus the NCCL gather across GPUs when more than one is visible."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from scenarios import GpuAdapter, assert_same_bits, load_golden, replay_default_scene, replay_forest_scene
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
    assert "MULTIGPU OK" in r.stdout and "MULTIGPU PEER OK" in r.stdout


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
