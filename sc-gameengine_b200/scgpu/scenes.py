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


# ---- lane graphs and on-rails agents for the traffic producer ---------------------------------

def _quant(v, scale):
    """quantPos / quantDir, src/engine/traffic/sc_traffic_lanes.cpp:34-44: round half away from zero"""
    s = np.float32(v) * np.float32(scale)
    return int(np.floor(np.float32(s + np.float32(0.5 if s >= 0 else -0.5))))


def lane_grid(nx, nz, sector=SECTOR_SIZE, lane_width=3.5, speed=12.0, x0=0, z0=0):
    """The lane graph TrafficLaneGraph::buildProceduralForSector (src/engine/traffic/sc_traffic_lanes.cpp:171-237)
    builds for the sectors [x0, x0+nx) x [z0, z0+nz), visited x-major: per sector one east/west and one north/south
    road of two opposite lanes, nodes shared between neighbouring sectors through the quantised (pos, dir) key of
    addNode (:70-97), so that lanes continue across sector borders. Returns the flat arrays of ScGpuLaneGraph."""
    f = np.float32
    nodes, key2node, conns, segs = [], {}, [], []

    def add_node(pos, d):
        k = (_quant(pos[0], 100), _quant(pos[1], 100), _quant(pos[2], 100), _quant(d[0], 1000), _quant(d[1], 1000), _quant(d[2], 1000))
        if k not in key2node:
            key2node[k] = len(nodes)
            nodes.append((f(pos[0]), f(pos[1]), f(pos[2])))
            conns.append([])
        return key2node[k]

    def add_segment(a, b):
        d = [f(nodes[b][i] - nodes[a][i]) for i in range(3)]
        ln = f(np.sqrt(f(f(f(d[0] * d[0]) + f(d[1] * d[1])) + f(d[2] * d[2]))))
        inv = f(f(1.0) / ln)
        conns[a].append(len(segs))
        segs.append((a, b, f(d[0] * inv), f(d[1] * inv), f(d[2] * inv), ln))

    off = f(f(lane_width) * f(0.5))
    for sx in range(x0, x0 + nx):
        for sz in range(z0, z0 + nz):
            mnx, mxx = f(f(sx) * f(sector)), f(f(sx + 1) * f(sector))
            mnz, mxz = f(f(sz) * f(sector)), f(f(sz + 1) * f(sector))
            cx, cz = f(f(mnx + mxx) * f(0.5)), f(f(mnz + mxz) * f(0.5))
            add_segment(add_node((mnx, 0, f(cz - off)), (1, 0, 0)), add_node((mxx, 0, f(cz - off)), (1, 0, 0)))
            add_segment(add_node((mxx, 0, f(cz + off)), (-1, 0, 0)), add_node((mnx, 0, f(cz + off)), (-1, 0, 0)))
            add_segment(add_node((f(cx + off), 0, mnz), (0, 0, 1)), add_node((f(cx + off), 0, mxz), (0, 0, 1)))
            add_segment(add_node((f(cx - off), 0, mxz), (0, 0, -1)), add_node((f(cx - off), 0, mnz), (0, 0, -1)))
    off_arr = np.zeros(len(nodes) + 1, np.uint32)
    off_arr[1:] = np.cumsum([len(c) for c in conns])
    sg = np.array(segs, np.float64).reshape(-1, 6)
    return dict(node_pos=np.array(nodes, np.float32).reshape(-1, 3), node_speed=np.full(len(nodes), speed, np.float32),
                conn_offset=off_arr, conn=np.array([s for c in conns for s in c], np.uint32),
                seg_nodes=np.ascontiguousarray(sg[:, :2].astype(np.uint32)), seg_dir=np.ascontiguousarray(sg[:, 2:5].astype(np.float32)),
                seg_len=np.ascontiguousarray(sg[:, 5].astype(np.float32)), seg_active=np.ones(len(segs), np.uint8),
                default_speed=np.float32(speed))


def lane_random(n_nodes, n_segs, seed=7, extent=400.0, hostile=True):
    """An arbitrary lane graph (curved roads, junctions with several exits, dead ends): random nodes, random directed
    segments with dir = (b - a) / |b - a|. With hostile=True some segments are inactive, some shorter than the 1e-5
    cut-off, some connection entries point past the segment array (skipped like sc_traffic_lanes.cpp:156-157) and
    speed limits include 0 and negative values."""
    rng = np.random.default_rng(seed)
    pos = (rng.random((n_nodes, 3), np.float32) * np.float32(extent)).astype(np.float32)
    pos[:, 1] = (rng.random(n_nodes, np.float32) * 3).astype(np.float32)
    a = rng.integers(0, n_nodes, n_segs).astype(np.uint32)
    b = rng.integers(0, n_nodes, n_segs).astype(np.uint32)
    d = pos[b] - pos[a]
    ln = np.sqrt((d * d).sum(1, dtype=np.float32)).astype(np.float32)
    sdir = np.where(ln[:, None] > 1e-6, d / np.maximum(ln, np.float32(1e-30))[:, None], np.float32([0, 0, 1])).astype(np.float32)
    active = np.ones(n_segs, np.uint8)
    speed = (4 + rng.random(n_nodes) * 20).astype(np.float32)
    conns = [[] for _ in range(n_nodes)]
    for s in range(n_segs):
        conns[a[s]].append(s)
    if hostile:
        active[rng.random(n_segs) < 0.08] = 0
        ln[rng.random(n_segs) < 0.03] = np.float32(5e-6)
        speed[rng.random(n_nodes) < 0.05] = 0
        speed[rng.random(n_nodes) < 0.03] = -3
        for c in conns:
            if rng.random() < 0.05:
                c.insert(int(rng.integers(0, len(c) + 1)), n_segs + int(rng.integers(0, 5)))
    off = np.zeros(n_nodes + 1, np.uint32)
    off[1:] = np.cumsum([len(c) for c in conns])
    return dict(node_pos=pos, node_speed=speed, conn_offset=off, conn=np.array([s for c in conns for s in c], np.uint32),
                seg_nodes=np.ascontiguousarray(np.stack([a, b], 1)), seg_dir=np.ascontiguousarray(sdir), seg_len=ln,
                seg_active=active, default_speed=np.float32(12.0))


def traffic_agents(graph, n, seed=11, hostile=True):
    """n on-rails agents placed on random segments of `graph` (what sc_traffic_spawner.cpp:283-316 produces): returns
    (agents dict lane/s/speed/look, trs9 [n, 9]). hostile adds agents without a lane, with lane ids past the array,
    off-lane positions and s beyond the segment end."""
    rng = np.random.default_rng(seed)
    ns = len(graph["seg_len"])
    lane = rng.integers(0, ns, n).astype(np.uint32)
    s = (rng.random(n, np.float32) * graph["seg_len"][lane]).astype(np.float32)
    a = graph["node_pos"][graph["seg_nodes"][lane, 0]]
    pos = (a + graph["seg_dir"][lane] * s[:, None]).astype(np.float32)
    pos[:, 1] = np.float32(0.9)
    yaw = np.arctan2(graph["seg_dir"][lane, 0], graph["seg_dir"][lane, 2]).astype(np.float32)
    speed = (rng.random(n, np.float32) * 14).astype(np.float32)
    look = np.full(n, 12.0, np.float32)
    if hostile:
        lane[rng.random(n) < 0.05] = 0xFFFFFFFF
        lane[rng.random(n) < 0.02] = ns + 3
        far = rng.random(n) < 0.05
        pos[far] += (rng.normal(size=(int(far.sum()), 3)) * 30).astype(np.float32)
        s[rng.random(n) < 0.03] += np.float32(500.0)
        look[rng.random(n) < 0.1] = np.float32(0.0)
        look[rng.random(n) < 0.1] = np.float32(150.0)
    trs = np.zeros((n, 9), np.float32)
    trs[:, 0:3] = pos
    trs[:, 3] = (rng.random(n) * 0.1).astype(np.float32)  # a vehicle that was tilted by the physics tier before
    trs[:, 4] = yaw
    trs[:, 6:9] = np.float32([1.8, 1.4, 4.2])
    return dict(lane=lane, s=s, speed=speed, look=look), trs
