"""Synthetic city scenes for the parity tests and bench.py.

Instance placement follows the reference's procedural sector generator
(/root/reference/src/engine/world/sc_world_partition.cpp:34-57 mix32/rand01, :105-169 generateSectorSpawnsStatic):
64 m sectors, props uniform inside the sector with a 1 m pad, y = sy/2, yaw in [0, 2pi), scale x,z in [0.4, 1.9],
y in [0.5, 3.2], unit-cube bounds. Everything is derived from one 32-bit seed, so CPU and GPU see the same scene.
This is harness code (numpy): it only produces inputs; no hot-path arithmetic happens here.
"""
from __future__ import annotations

import numpy as np

SECTOR_SIZE = 64.0
NO_PARENT = -1
_U = np.uint32


def mix32(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    x ^= x >> _U(16)
    x *= _U(0x7FEB352D)
    x ^= x >> _U(15)
    x *= _U(0x846CA68B)
    x ^= x >> _U(16)
    return x


def hash_coord_seed(seed, cx, cz):
    h = np.full(cx.shape, seed, dtype=np.uint32)
    h ^= mix32(cx.astype(np.int32).view(np.uint32) * _U(73856093))
    h ^= mix32(cz.astype(np.int32).view(np.uint32) * _U(19349663))
    return mix32(h + _U(0x9E3779B9))


def rand01(state):
    """state: uint32 array, advanced in place; returns float32 in [0,1]"""
    state[...] = mix32(state + _U(0x6D2B79F5))
    return (state & _U(0x00FFFFFF)).astype(np.float32) / np.float32(16777215.0)


def _lerp(a, b, t):
    a = np.float32(a)
    b = np.float32(b)
    return (a + (b - a) * t).astype(np.float32)


def city_props(n, seed=424242, props_min=18, props_max=34):
    """n props laid out sector by sector on a square grid of sectors centred on the origin sector.
    Returns dict(trs9 [n,9] f32, sector [n,2] i32, mesh_mat [n,2] u32, grid side)."""
    mean = (props_min + props_max) / 2.0
    n_sectors = int(np.ceil(n / mean * 1.02)) + 4
    side = int(np.ceil(np.sqrt(n_sectors)))
    while True:
        gx, gz = np.meshgrid(np.arange(side, dtype=np.int32), np.arange(side, dtype=np.int32), indexing="xy")
        cx = (gx.ravel() - side // 2).astype(np.int32)
        cz = (gz.ravel() - side // 2).astype(np.int32)
        rng = hash_coord_seed(seed, cx, cz)
        count_range = props_max - props_min + 1
        prop_count = (props_min + (mix32(rng) % _U(count_range))).astype(np.int64)
        if prop_count.sum() >= n:
            break
        side += 1
    ns = cx.shape[0]
    kmax = props_max
    size = np.float32(SECTOR_SIZE)
    min_x = cx.astype(np.float32) * size
    min_z = cz.astype(np.float32) * size
    pad = np.float32(1.0)
    trs = np.zeros((ns, kmax, 9), np.float32)
    mm = np.zeros((ns, kmax, 2), np.uint32)
    state = rng.copy()
    for k in range(kmax):
        x = (min_x + pad) + ((min_x + size - pad) - (min_x + pad)) * rand01(state)
        z = (min_z + pad) + ((min_z + size - pad) - (min_z + pad)) * rand01(state)
        sx = _lerp(0.4, 1.9, rand01(state))
        sy = _lerp(0.5, 3.2, rand01(state))
        sz = _lerp(0.4, 1.9, rand01(state))
        yaw = rand01(state) * np.float32(3.1415926535 * 2.0)
        m = rand01(state)
        mesh = rand01(state)
        trs[:, k, 0] = x
        trs[:, k, 1] = sy * np.float32(0.5)
        trs[:, k, 2] = z
        trs[:, k, 4] = yaw
        trs[:, k, 6] = sx
        trs[:, k, 7] = sy
        trs[:, k, 8] = sz
        mm[:, k, 0] = np.where(mesh < np.float32(0.90), 1, 2)
        mm[:, k, 1] = np.where(m < np.float32(0.40), 1, np.where(m < np.float32(0.80), 2, 3))
    keep = np.arange(kmax)[None, :] < prop_count[:, None]
    trs9 = trs[keep][:n]
    mesh_mat = mm[keep][:n]
    sec = np.stack([np.repeat(cx, kmax).reshape(ns, kmax)[keep][:n], np.repeat(cz, kmax).reshape(ns, kmax)[keep][:n]], axis=1)
    return dict(trs9=np.ascontiguousarray(trs9), sector=np.ascontiguousarray(sec.astype(np.int32)),
                mesh_mat=np.ascontiguousarray(mesh_mat), side=side)


def city_flat(n, seed=424242):
    """Config 2: n flat instances (no parenting), unit-cube bounds, all with RenderMesh."""
    p = city_props(n, seed)
    return dict(
        n=n,
        trs9=p["trs9"],
        parent=np.full(n, NO_PARENT, np.int64),
        aabb6=np.tile(np.array([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5], np.float32), (n, 1)),
        mesh_mat=p["mesh_mat"],
        flags=np.full(n, 3, np.uint32),
        sector=p["sector"],
        side=p["side"],
    )


# group templates: (parent index inside the group or -1, local offset, local scale, rotates-about axis)
_VEHICLE = [  # vehicle root -> body -> 4 wheels -> 1 attachment per wheel  (depth 4, 10 nodes)
    (-1, (0, 0, 0), (1, 1, 1)),
    (0, (0, 0.6, 0), (1.8, 0.9, 4.2)),
    (1, (-0.9, -0.4, 1.3), (0.35, 0.35, 0.2)),
    (1, (0.9, -0.4, 1.3), (0.35, 0.35, 0.2)),
    (1, (-0.9, -0.4, -1.3), (0.35, 0.35, 0.2)),
    (1, (0.9, -0.4, -1.3), (0.35, 0.35, 0.2)),
    (2, (0, 0, 0.15), (0.5, 0.5, 0.5)),
    (3, (0, 0, 0.15), (0.5, 0.5, 0.5)),
    (4, (0, 0, 0.15), (0.5, 0.5, 0.5)),
    (5, (0, 0, 0.15), (0.5, 0.5, 0.5)),
]
_PED = [  # ped root -> torso -> limb -> prop  (depth 4, 4 nodes)
    (-1, (0, 0, 0), (1, 1, 1)),
    (0, (0, 1.1, 0), (0.5, 0.7, 0.3)),
    (1, (0.4, 0.2, 0), (0.2, 0.9, 0.2)),
    (2, (0, -0.6, 0.1), (0.6, 0.6, 0.6)),
]


def city_hier(n, seed=424242):
    """Config 3: n instances in depth-4 groups (vehicles + wheels, peds + attachments), spawned group by group.
    Group roots are placed like city props; children carry small local offsets and rotations about X/Y."""
    n_groups_est = int(np.ceil(n / (4.0 + 6.0 * float(__import__('os').environ.get('SCGPU_VEHICLE_FRACTION', '0.5'))) * 1.2)) + 16
    roots = city_props(n_groups_est, seed)
    rng = np.random.default_rng(seed)
    import os as _os
    is_vehicle = rng.random(n_groups_est) < float(_os.environ.get('SCGPU_VEHICLE_FRACTION', '0.5'))
    sizes = np.where(is_vehicle, len(_VEHICLE), len(_PED))
    starts = np.concatenate([[0], np.cumsum(sizes)])
    total = int(starts[-1])
    assert total >= n, (total, n)
    trs9 = np.zeros((total, 9), np.float32)
    parent = np.full(total, NO_PARENT, np.int64)
    sector = np.zeros((total, 2), np.int32)
    mesh_mat = np.zeros((total, 2), np.uint32)
    for tmpl, sel in ((_VEHICLE, is_vehicle), (_PED, ~is_vehicle)):
        g = np.nonzero(sel)[0]
        base = starts[g]
        spin = rng.random((g.shape[0], len(tmpl))).astype(np.float32) * np.float32(6.2831853)
        for k, (pk, off, scl) in enumerate(tmpl):
            idx = base + k
            if pk < 0:
                trs9[idx] = roots["trs9"][g]
                trs9[idx, 1] = 0.0
                trs9[idx, 6:9] = 1.0
            else:
                trs9[idx, 0:3] = np.array(off, np.float32)
                trs9[idx, 6:9] = np.array(scl, np.float32)
                # wheels and limbs rotate about X, attachments about Y
                trs9[idx, 3 if (k % 2 == 0) else 4] = spin[:, k]
                parent[idx] = base + pk
            sector[idx] = roots["sector"][g]
            mesh_mat[idx, 0] = 1 + (k % 3)
            mesh_mat[idx, 1] = 1 + (k % 4)
    trs9, parent, sector, mesh_mat = trs9[:n], parent[:n], sector[:n], mesh_mat[:n]
    parent = np.where(parent >= n, NO_PARENT, parent)  # a group cut by the truncation keeps no dangling index
    flags = np.full(n, 3, np.uint32)
    flags[parent == NO_PARENT] = 1  # group roots are transform-only nodes (no RenderMesh), like the sandbox's Root
    return dict(n=n, trs9=np.ascontiguousarray(trs9), parent=parent,
                aabb6=np.tile(np.array([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5], np.float32), (n, 1)),
                mesh_mat=np.ascontiguousarray(mesh_mat), flags=flags, sector=np.ascontiguousarray(sector),
                side=roots["side"])


def parent_handles(parent_index, entity):
    """batch-relative parent indices (-1 = none) -> entity handles (0xFFFFFFFF = none)"""
    parent_index = np.asarray(parent_index)
    out = np.full(parent_index.shape, 0xFFFFFFFF, np.uint32)
    m = parent_index >= 0
    out[m] = np.asarray(entity, np.uint32)[parent_index[m]]
    return out


# ---- views ------------------------------------------------------------------------------------------------

def _mat_trs(pos, rot):
    cx, sx = np.cos(rot[0]), np.sin(rot[0])
    cy, sy = np.cos(rot[1]), np.sin(rot[1])
    cz, sz = np.cos(rot[2]), np.sin(rot[2])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    m = np.eye(4)
    m[:3, :3] = rz @ ry @ rx
    m[:3, 3] = pos
    return m


def perspective_view_proj(pos, rot, fov_deg=60.0, aspect=1280.0 / 720.0, near=0.1, far=1000.0):
    """Column-major float32[16] of P * inverse(camWorld), the shape CameraSystem produces
    (src/core/src/sc_ecs.cpp:213-272; RH, depth 0..1, flipY)."""
    f = 1.0 / np.tan(np.radians(fov_deg) * 0.5)
    p = np.zeros((4, 4))
    p[0, 0] = f / aspect
    p[1, 1] = -f
    p[2, 2] = far / (near - far)
    p[2, 3] = (far * near) / (near - far)
    p[3, 2] = -1.0
    view = np.linalg.inv(_mat_trs(np.asarray(pos, float), np.asarray(rot, float)))
    return np.ascontiguousarray((p @ view).T.astype(np.float32).ravel())


def ortho_view_proj(center, half_extent, light_dir=(0.35, -0.85, 0.4), depth=2000.0):
    """Orthographic shadow-cascade matrix (the reference has no ortho helper)."""
    d = np.asarray(light_dir, float)
    d /= np.linalg.norm(d)
    up = np.array([0.0, 1.0, 0.0])
    r = np.cross(d, up)
    r /= np.linalg.norm(r)
    u = np.cross(r, d)
    eye = np.asarray(center, float) - d * depth * 0.5
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = r, u, -d
    view[:3, 3] = -view[:3, :3] @ eye
    o = np.eye(4)
    o[0, 0] = 1.0 / half_extent
    o[1, 1] = 1.0 / half_extent
    o[2, 2] = -1.0 / depth
    return np.ascontiguousarray((o @ view).T.astype(np.float32).ravel())


def standard_views(n_views, center=(32.0, 6.0, 44.0), yaw=3.14159265, cascades=(25.0, 75.0, 200.0, 600.0)):
    """main perspective camera (+ up to 4 cascades, then extra perspective cameras looking elsewhere)"""
    views = [perspective_view_proj(center, (0.0, yaw, 0.0))]
    for h in cascades:
        if len(views) >= n_views:
            break
        views.append(ortho_view_proj((center[0], 0.0, center[2]), h))
    k = 1
    while len(views) < n_views:
        views.append(perspective_view_proj(center, (0.0, yaw + 1.3 * k, 0.0)))
        k += 1
    return np.stack(views[:n_views]).astype(np.float32)


def shard_by_sector(sector, n_ranks):
    """World-cell sharding: contiguous blocks of sectors in row-major (z, x) order, balanced by
    instance count. Returns rank id per instance. Instances of one sector never split across ranks."""
    sector = np.asarray(sector)
    key = (sector[:, 1].astype(np.int64) << 32) + (sector[:, 0].astype(np.int64) & 0xFFFFFFFF)
    uniq, inv, counts = np.unique(key, return_inverse=True, return_counts=True)
    cum = np.cumsum(counts)
    total = cum[-1]
    # sector k goes to the rank whose share contains its midpoint
    mid = cum - counts / 2.0
    sec_rank = np.minimum((mid * n_ranks / total).astype(np.int64), n_ranks - 1)
    return sec_rank[inv].astype(np.int32)
