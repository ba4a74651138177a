"""scgpu — ctypes binding of libscgpu.so (include/scgpu.h), used by tests/ and bench.py.

The product is the C ABI; this module is a thin, allocation-free-as-possible mirror of it for Python harnesses.
It never falls back to a CPU path: if the CUDA library is missing or no B200 is present it raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE.parent / "libscgpu.so"

MAX_VIEWS = 8
INVALID_ENTITY = 0xFFFFFFFF
HAS_BOUNDS = 1
HAS_MESH = 2
UPDATE_FREEZE_CULLING = 1
UPDATE_SKIP_TRANSFORM = 2
UPDATE_CULLED_LISTS = 4
COMM_ID_BYTES = 128


class ScGpuError(RuntimeError):
    pass


class SceneDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("max_instances", C.c_uint32),
        ("max_entity_index", C.c_uint32),
        ("max_views", C.c_uint32),
        ("flags", C.c_uint32),
        ("stream", C.c_void_p),
    ]


class Counts(C.Structure):
    _fields_ = [
        ("transforms", C.c_uint32),
        ("renderablesTotal", C.c_uint32),
        ("visible", C.c_uint32 * MAX_VIEWS),
        ("culled", C.c_uint32 * MAX_VIEWS),
        ("recomputed", C.c_uint32),
        ("slowWindows", C.c_uint32),
        ("extent", C.c_uint32),
    ]


class DeviceViews(C.Structure):
    _fields_ = [
        ("visibleEntity", C.c_void_p * MAX_VIEWS),
        ("visibleSlot", C.c_void_p * MAX_VIEWS),
        ("visibleCount", C.c_void_p),
        ("worldCol", C.c_void_p * 4),
        ("entity", C.c_void_p),
        ("count", C.c_uint32),
        ("extent", C.c_uint32),
        ("rank", C.c_void_p),
        ("perm", C.c_void_p),
    ]


DRAW_ITEM_DTYPE = np.dtype(
    [("entity", "<u4"), ("meshId", "<u4"), ("materialId", "<u4"), ("_pad", "<u4"), ("model", "<f4", (16,))]
)
assert DRAW_ITEM_DTYPE.itemsize == 80
class AssetBinding(C.Structure):
    _fields_ = [("assetId", C.c_uint64), ("handle", C.c_uint32), ("_pad", C.c_uint32)]


class AssetTable(C.Structure):
    _fields_ = [("meshes", C.POINTER(AssetBinding)), ("nMeshes", C.c_uint32), ("defaultMesh", C.c_uint32),
                ("materials", C.POINTER(AssetBinding)), ("nMaterials", C.c_uint32), ("defaultMaterial", C.c_uint32)]


def make_asset_table(meshes: dict, default_mesh: int, materials: dict, default_material: int) -> "AssetTable":
    """{assetId: handle} dicts -> ScGpuAssetTable (keeps the arrays alive on the returned object)"""
    m = (AssetBinding * max(len(meshes), 1))(*[AssetBinding(int(k), int(h), 0) for k, h in meshes.items()])
    t = (AssetBinding * max(len(materials), 1))(*[AssetBinding(int(k), int(h), 0) for k, h in materials.items()])
    tab = AssetTable(m, len(meshes), default_mesh, t, len(materials), default_material)
    tab._keep = (m, t)
    return tab


class SectorGen(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("sectorSizeMeters", C.c_float), ("seed", C.c_uint32),
                ("propsPerSectorMin", C.c_uint32), ("propsPerSectorMax", C.c_uint32), ("includeGroundPlane", C.c_uint32),
                ("meshCube", C.c_uint32), ("meshTriangle", C.c_uint32),
                ("matUnlit", C.c_uint32), ("matChecker", C.c_uint32), ("matTest", C.c_uint32)]

    def __init__(self, **kw):
        super().__init__(struct_size=C.sizeof(SectorGen), **kw)


class LaneGraph(C.Structure):
    """ScGpuLaneGraph (include/scgpu.h)"""
    _fields_ = [("struct_size", C.c_uint32), ("nNodes", C.c_uint32), ("nSegments", C.c_uint32), ("nConnections", C.c_uint32),
                ("nodePos", C.c_void_p), ("nodeSpeedLimit", C.c_void_p), ("nodeConnOffset", C.c_void_p), ("nodeConn", C.c_void_p),
                ("segNodes", C.c_void_p), ("segDir", C.c_void_p), ("segLength", C.c_void_p), ("segActive", C.c_void_p),
                ("defaultSpeedLimit", C.c_float)]


class TrafficStep(C.Structure):
    """ScGpuTrafficStep (include/scgpu.h)"""
    _fields_ = [("struct_size", C.c_uint32), ("dt", C.c_float), ("hasDebug", C.c_uint32), ("lookAheadDist", C.c_float),
                ("speedMultiplier", C.c_float), ("obstacleBrake", C.c_void_p), ("skip", C.c_void_p)]


EDITOR_DRAW_DTYPE = np.dtype([("mesh", "<u8"), ("material", "<u8"), ("model", "<f4", (16,)), ("flags", "<u4"), ("_pad", "<u4")])
assert EDITOR_DRAW_DTYPE.itemsize == 88
DRAW_RUN_DTYPE = np.dtype([("pipelineId", "<u4"), ("materialId", "<u4"), ("meshId", "<u4"), ("first", "<u4"), ("count", "<u4")])
assert DRAW_RUN_DTYPE.itemsize == 20

_lib = None

# every symbol include/scgpu.h declares: (name, restype, argtypes)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p
SYMBOLS = {
    "scgpuGetApiVersion": (C.c_uint32, []),
    "scgpuCreate": (_vp, [C.POINTER(SceneDesc)]),
    "scgpuDestroy": (None, [_vp]),
    "scgpuLastError": (C.c_char_p, [_vp]),
    "scgpuSpawn": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scgpuDespawn": (C.c_int, [_vp, C.c_uint32, _vp]),
    "scgpuBuildEditorDraws": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_uint32, _u32p]),
    "scgpuSectorFileInfo": (C.c_int, [_vp, C.c_size_t, _vp, _u32p, _u32p]),
    "scgpuSpawnSectorFile": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_uint32, C.POINTER(AssetTable)]),
    "scgpuSectorSpawnCount": (C.c_uint32, [C.POINTER(SectorGen), C.c_int32, C.c_int32]),
    "scgpuSpawnSectors": (C.c_int, [_vp, C.POINTER(SectorGen), C.c_uint32, _vp, _vp, C.c_uint32]),
    "scgpuSetLocal": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuSetLocalPosRot": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuSetLocalPosition": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuSetLocalRange": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "scgpuSetRender": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "scgpuSetParent": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuMarkDirty": (C.c_int, [_vp, C.c_uint32, _vp]),
    "scgpuSetLocalDevice": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuMarkAllDirty": (C.c_int, [_vp]),
    "scgpuSetViews": (C.c_int, [_vp, C.c_uint32, _vp]),
    "scgpuSetViewPlanes": (C.c_int, [_vp, C.c_uint32, _vp]),
    "scgpuGetViewPlanes": (C.c_int, [_vp, C.c_uint32, _vp]),
    "scgpuUpdate": (C.c_int, [_vp, C.c_uint32]),
    "scgpuSynchronize": (C.c_int, [_vp]),
    "scgpuGetCounts": (C.c_int, [_vp, C.POINTER(Counts)]),
    "scgpuReadVisible": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, _u32p]),
    "scgpuReadCulled": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, _u32p]),
    "scgpuReadDrawItems": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32, _u32p, _u32p]),
    "scgpuReadWorld": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuReadDenseEntities": (C.c_int, [_vp, _vp, C.c_uint32, _u32p]),
    "scgpuReadParents": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuGetDeviceViews": (C.c_int, [_vp, C.POINTER(DeviceViews)]),
    "scgpuGetStream": (_vp, [_vp]),
    "scgpuBuildDrawItemsDevice": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.POINTER(_vp), _u32p, _u32p]),
    "scgpuBuildSortedDraws": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp), _u32p,
                                        C.POINTER(_vp), _u32p]),
    "scgpuReadSortedDraws": (C.c_int, [_vp, _vp, C.c_uint32, _vp, C.c_uint32]),
    "scgpuTrafficSetLanes": (C.c_int, [_vp, C.POINTER(LaneGraph)]),
    "scgpuTrafficSetLaneActive": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuTrafficSetAgents": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "scgpuTrafficAdvance": (C.c_int, [_vp, C.POINTER(TrafficStep), _u32p]),
    "scgpuTrafficReadAgents": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _u32p]),
    "scgpuReadLocal": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "scgpuCommGetUniqueId": (C.c_int, [_vp]),
    "scgpuCommInit": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "scgpuCommEnablePeerGather": (C.c_int, [_vp, C.c_uint32, C.c_uint32]),
    "scgpuGatherVisible": (C.c_int, [_vp, C.c_uint32]),
    "scgpuGetGatheredCounts": (C.c_int, [_vp, _vp, C.c_uint32]),
    "scgpuReadGatheredVisible": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, _u32p]),
    "scgpuKernelLaunchCount": (C.c_uint64, [_vp]),
    "scgpuLastUpdateTimings": (C.c_int, [_vp, _f32p, _f32p]),
    "scgpuEnableTimings": (C.c_int, [_vp, C.c_int]),
    "scgpuReadUpdateTimings": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _u32p]),
}


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Loads libscgpu.so (built in-tree by sc-gameengine_b200/Makefile). Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # SCGPU_LIB: A/B-testing hook for kernel build variants (bench only)
    p = Path(path) if path else Path(os.environ.get("SCGPU_LIB", LIB_PATH))
    if not p.exists():
        raise ScGpuError(f"{p} not found: build it with `make -C sc-gameengine_b200` (there is no CPU fallback)")
    lib = C.CDLL(str(p))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _arr(a, dtype, shape_last=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape_last is not None and (a.ndim != 2 or a.shape[1] != shape_last):
        a = a.reshape(-1, shape_last)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Scene:
    """One scgpu context (one GPU). Mirrors include/scgpu.h one to one."""

    def __init__(self, max_instances: int, max_views: int = 1, device: int = 0, max_entity_index: int = 0,
                 stream: int | None = None):
        self.lib = load_library()
        d = SceneDesc(C.sizeof(SceneDesc), device, max_instances, max_entity_index, max_views, 0, stream)
        self.ctx = self.lib.scgpuCreate(C.byref(d))
        if not self.ctx:
            raise ScGpuError("scgpuCreate failed: " + self.lib.scgpuLastError(None).decode())
        self.max_instances = max_instances
        self.max_views = max_views
        self.n_views = 0

    # -- helpers
    def _ck(self, ok, what):
        if not ok:
            raise ScGpuError(f"{what}: {self.lib.scgpuLastError(self.ctx).decode()}")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.scgpuDestroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- deltas
    def spawn(self, entity, trs9, parent=None, aabb6=None, mesh_mat=None, flags=None):
        e = _arr(entity, np.uint32)
        t = _arr(trs9, np.float32, 9)
        p = _arr(parent, np.uint32)
        b = _arr(aabb6, np.float32, 6)
        m = _arr(mesh_mat, np.uint32, 2)
        f = _arr(flags, np.uint32)
        n = e.shape[0]
        assert t.shape[0] == n
        self._ck(self.lib.scgpuSpawn(self.ctx, n, _ptr(e), _ptr(p), _ptr(t), _ptr(b), _ptr(m), _ptr(f)), "scgpuSpawn")

    def spawn_sectors(self, gen: "SectorGen", coord_xz, entity):
        """procedural sectors generated on the device; entity = handles of all sectors back to back"""
        cxz = _arr(coord_xz, np.int32).reshape(-1, 2)
        e = _arr(entity, np.uint32)
        self._ck(self.lib.scgpuSpawnSectors(self.ctx, C.byref(gen), cxz.shape[0], _ptr(cxz), _ptr(e), e.shape[0]), "scgpuSpawnSectors")

    def spawn_sector_file(self, file_bytes, entity, assets: "AssetTable"):
        """a .scsector file image; its INST chunk is unpacked into the SoA on the device"""
        raw = np.ascontiguousarray(np.frombuffer(bytes(file_bytes), np.uint8))
        e = _arr(entity, np.uint32)
        self._ck(self.lib.scgpuSpawnSectorFile(self.ctx, _ptr(raw), raw.shape[0], _ptr(e), e.shape[0], C.byref(assets)),
                 "scgpuSpawnSectorFile")

    def editor_draws(self, trs9, mesh, material):
        """the world editor's BuildDrawItems on the device"""
        t = _arr(trs9, np.float32).reshape(-1, 9)
        m, a = _arr(mesh, np.uint64), _arr(material, np.uint64)
        out = np.zeros(t.shape[0], EDITOR_DRAW_DTYPE)
        n = C.c_uint32(0)
        self._ck(self.lib.scgpuBuildEditorDraws(self.ctx, t.shape[0], _ptr(t), _ptr(m), _ptr(a), _ptr(out), out.shape[0], C.byref(n)),
                 "scgpuBuildEditorDraws")
        return out[: n.value]

    # -- traffic on rails
    def traffic_set_lanes(self, node_pos, node_speed, conn_offset, conn, seg_nodes, seg_dir, seg_len, seg_active=None,
                          default_speed=12.0):
        a = [_arr(node_pos, np.float32), _arr(node_speed, np.float32), _arr(conn_offset, np.uint32), _arr(conn, np.uint32),
             _arr(seg_nodes, np.uint32), _arr(seg_dir, np.float32), _arr(seg_len, np.float32), _arr(seg_active, np.uint8)]
        g = LaneGraph(C.sizeof(LaneGraph), a[1].shape[0], a[6].shape[0], a[3].shape[0], *[_ptr(x) for x in a], float(default_speed))
        self._ck(self.lib.scgpuTrafficSetLanes(self.ctx, C.byref(g)), "scgpuTrafficSetLanes")

    def traffic_set_lane_active(self, segment, active):
        sg, ac = _arr(segment, np.uint32), _arr(active, np.uint8)
        self._ck(self.lib.scgpuTrafficSetLaneActive(self.ctx, sg.shape[0], _ptr(sg), _ptr(ac)), "scgpuTrafficSetLaneActive")

    def traffic_set_agents(self, entity, lane, s, speed, look):
        e = _arr(entity, np.uint32)
        self._ck(self.lib.scgpuTrafficSetAgents(self.ctx, e.shape[0], _ptr(e), _ptr(_arr(lane, np.uint32)), _ptr(_arr(s, np.float32)),
                                                _ptr(_arr(speed, np.float32)), _ptr(_arr(look, np.float32))), "scgpuTrafficSetAgents")
        self._n_agents = e.shape[0]

    def traffic_advance(self, dt, brake=None, skip=None, debug=None, want_moved=True):
        """one TrafficAISystem pass over the on-rails agents; debug = (lookAheadDist, speedMultiplier) or None"""
        b, k = _arr(brake, np.float32), _arr(skip, np.uint8)
        st = TrafficStep(C.sizeof(TrafficStep), dt, 1 if debug else 0, debug[0] if debug else 0.0, debug[1] if debug else 0.0,
                         _ptr(b), _ptr(k))
        moved = C.c_uint32(0)
        self._ck(self.lib.scgpuTrafficAdvance(self.ctx, C.byref(st), C.byref(moved) if want_moved else None), "scgpuTrafficAdvance")
        return moved.value

    def traffic_read_agents(self):
        n = getattr(self, "_n_agents", 0)
        lane, s, v, look = np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        cnt = C.c_uint32(0)
        self._ck(self.lib.scgpuTrafficReadAgents(self.ctx, n, _ptr(lane), _ptr(s), _ptr(v), _ptr(look), C.byref(cnt)),
                 "scgpuTrafficReadAgents")
        assert cnt.value == n
        return lane, s, v, look

    def read_local(self, entity):
        e = _arr(entity, np.uint32)
        out = np.zeros((e.shape[0], 9), np.float32)
        self._ck(self.lib.scgpuReadLocal(self.ctx, e.shape[0], _ptr(e), _ptr(out)), "scgpuReadLocal")
        return out

    def despawn(self, entity):
        e = _arr(entity, np.uint32)
        self._ck(self.lib.scgpuDespawn(self.ctx, e.shape[0], _ptr(e)), "scgpuDespawn")

    def set_local(self, entity, trs9):
        e = _arr(entity, np.uint32)
        t = _arr(trs9, np.float32, 9)
        self._ck(self.lib.scgpuSetLocal(self.ctx, e.shape[0], _ptr(e), _ptr(t)), "scgpuSetLocal")

    def set_local_pos_rot(self, entity, pos_rot6):
        e = _arr(entity, np.uint32)
        t = _arr(pos_rot6, np.float32, 6)
        self._ck(self.lib.scgpuSetLocalPosRot(self.ctx, e.shape[0], _ptr(e), _ptr(t)), "scgpuSetLocalPosRot")

    def set_local_position(self, entity, pos3):
        e = _arr(entity, np.uint32)
        t = _arr(pos3, np.float32, 3)
        self._ck(self.lib.scgpuSetLocalPosition(self.ctx, e.shape[0], _ptr(e), _ptr(t)), "scgpuSetLocalPosition")

    def set_local_range(self, first_dense, data, floats_per_instance):
        """fields of the Transforms at dense indices first_dense.. (pool order); 3 = pos, 6 = pos + rot, 9 = TRS"""
        t = _arr(data, np.float32, floats_per_instance)
        self._ck(self.lib.scgpuSetLocalRange(self.ctx, first_dense, t.shape[0], floats_per_instance, _ptr(t)), "scgpuSetLocalRange")

    def set_render(self, entity, mesh_mat=None, aabb6=None, flags=None):
        e = _arr(entity, np.uint32)
        self._ck(self.lib.scgpuSetRender(self.ctx, e.shape[0], _ptr(e), _ptr(_arr(mesh_mat, np.uint32, 2)),
                                         _ptr(_arr(aabb6, np.float32, 6)), _ptr(_arr(flags, np.uint32))), "scgpuSetRender")

    def set_local_device(self, n, d_entity_ptr, d_trs_ptr):
        self._ck(self.lib.scgpuSetLocalDevice(self.ctx, n, d_entity_ptr, d_trs_ptr), "scgpuSetLocalDevice")

    def set_parent(self, entity, parent):
        e = _arr(entity, np.uint32)
        p = _arr(parent, np.uint32)
        self._ck(self.lib.scgpuSetParent(self.ctx, e.shape[0], _ptr(e), _ptr(p)), "scgpuSetParent")

    def mark_dirty(self, entity):
        e = _arr(entity, np.uint32)
        self._ck(self.lib.scgpuMarkDirty(self.ctx, e.shape[0], _ptr(e)), "scgpuMarkDirty")

    def mark_all_dirty(self):
        self._ck(self.lib.scgpuMarkAllDirty(self.ctx), "scgpuMarkAllDirty")

    # -- views
    def set_views(self, view_proj):
        vp = _arr(view_proj, np.float32).reshape(-1, 16)
        self._ck(self.lib.scgpuSetViews(self.ctx, vp.shape[0], _ptr(vp)), "scgpuSetViews")
        self.n_views = vp.shape[0]

    def set_view_planes(self, planes):
        pl = _arr(planes, np.float32).reshape(-1, 24)
        self._ck(self.lib.scgpuSetViewPlanes(self.ctx, pl.shape[0], _ptr(pl)), "scgpuSetViewPlanes")
        self.n_views = pl.shape[0]

    def get_view_planes(self, view):
        out = np.zeros(24, np.float32)
        self._ck(self.lib.scgpuGetViewPlanes(self.ctx, view, _ptr(out)), "scgpuGetViewPlanes")
        return out.reshape(6, 4)

    # -- frame
    def update(self, flags: int = 0):
        self._ck(self.lib.scgpuUpdate(self.ctx, flags), "scgpuUpdate")

    def synchronize(self):
        self._ck(self.lib.scgpuSynchronize(self.ctx), "scgpuSynchronize")

    # -- results
    def counts(self) -> Counts:
        c = Counts()
        self._ck(self.lib.scgpuGetCounts(self.ctx, C.byref(c)), "scgpuGetCounts")
        return c

    def read_visible(self, view=0, out=None):
        n = C.c_uint32(0)
        self._ck(self.lib.scgpuReadVisible(self.ctx, view, None, 0, C.byref(n)), "scgpuReadVisible")
        if out is None:
            out = np.empty(n.value, np.uint32)
        self._ck(self.lib.scgpuReadVisible(self.ctx, view, _ptr(out), out.shape[0], C.byref(n)), "scgpuReadVisible")
        return out[: n.value]

    def read_culled(self, view=0):
        n = C.c_uint32(0)
        self._ck(self.lib.scgpuReadCulled(self.ctx, view, None, 0, C.byref(n)), "scgpuReadCulled")
        out = np.empty(n.value, np.uint32)
        self._ck(self.lib.scgpuReadCulled(self.ctx, view, _ptr(out), out.shape[0], C.byref(n)), "scgpuReadCulled")
        return out

    def read_draw_items(self, view=0, max_draws=0):
        e = C.c_uint32(0)
        d = C.c_uint32(0)
        self._ck(self.lib.scgpuReadDrawItems(self.ctx, view, max_draws, None, 0, C.byref(e), C.byref(d)), "scgpuReadDrawItems")
        out = np.zeros(e.value, DRAW_ITEM_DTYPE)
        self._ck(self.lib.scgpuReadDrawItems(self.ctx, view, max_draws, _ptr(out), out.shape[0], C.byref(e), C.byref(d)),
                 "scgpuReadDrawItems")
        return out, e.value, d.value

    def sorted_draws(self, view, material_pipeline, mesh_count, max_draws=0):
        """(items sorted by (pipeline, material, mesh), runs) for one view"""
        mp = _arr(material_pipeline, np.uint32)
        kept, runs = C.c_uint32(0), C.c_uint32(0)
        self._ck(self.lib.scgpuBuildSortedDraws(self.ctx, view, max_draws, _ptr(mp), mp.shape[0], mesh_count, None,
                                                C.byref(kept), None, C.byref(runs)), "scgpuBuildSortedDraws")
        items = np.zeros(kept.value, DRAW_ITEM_DTYPE)
        r = np.zeros(runs.value, DRAW_RUN_DTYPE)
        self._ck(self.lib.scgpuReadSortedDraws(self.ctx, _ptr(items), items.shape[0], _ptr(r), r.shape[0]), "scgpuReadSortedDraws")
        return items, r

    def read_world(self, entity):
        e = _arr(entity, np.uint32)
        out = np.zeros((e.shape[0], 16), np.float32)
        self._ck(self.lib.scgpuReadWorld(self.ctx, e.shape[0], _ptr(e), _ptr(out)), "scgpuReadWorld")
        return out

    def read_parents(self, entity):
        e = _arr(entity, np.uint32)
        out = np.zeros(e.shape[0], np.uint32)
        self._ck(self.lib.scgpuReadParents(self.ctx, e.shape[0], _ptr(e), _ptr(out)), "scgpuReadParents")
        return out

    def read_dense_entities(self):
        n = C.c_uint32(0)
        self._ck(self.lib.scgpuReadDenseEntities(self.ctx, None, 0, C.byref(n)), "scgpuReadDenseEntities")
        out = np.empty(n.value, np.uint32)
        self._ck(self.lib.scgpuReadDenseEntities(self.ctx, _ptr(out), out.shape[0], C.byref(n)), "scgpuReadDenseEntities")
        return out

    def device_views(self) -> DeviceViews:
        dv = DeviceViews()
        self._ck(self.lib.scgpuGetDeviceViews(self.ctx, C.byref(dv)), "scgpuGetDeviceViews")
        return dv

    def build_draw_items_device(self, view=0, max_draws=0):
        p = C.c_void_p(0)
        e = C.c_uint32(0)
        d = C.c_uint32(0)
        self._ck(self.lib.scgpuBuildDrawItemsDevice(self.ctx, view, max_draws, C.byref(p), C.byref(e), C.byref(d)),
                 "scgpuBuildDrawItemsDevice")
        return p.value, e.value, d.value

    @property
    def stream(self) -> int:
        return self.lib.scgpuGetStream(self.ctx) or 0

    @property
    def launches(self) -> int:
        return int(self.lib.scgpuKernelLaunchCount(self.ctx))

    def enable_timings(self, on=True, every=1):
        """every = n > 1: only every n-th update carries the event pairs (they cut the dependent-launch chain)"""
        self._ck(self.lib.scgpuEnableTimings(self.ctx, (max(1, int(every)) if on else 0)), "scgpuEnableTimings")

    def last_timings(self):
        k = C.c_float(0)
        u = C.c_float(0)
        self._ck(self.lib.scgpuLastUpdateTimings(self.ctx, C.byref(k), C.byref(u)), "scgpuLastUpdateTimings")
        return k.value, u.value

    def read_timings(self, cap=256):
        k = np.zeros(cap, np.float32)
        u = np.zeros(cap, np.float32)
        n = C.c_uint32(0)
        self._ck(self.lib.scgpuReadUpdateTimings(self.ctx, _ptr(k), _ptr(u), cap, C.byref(n)), "scgpuReadUpdateTimings")
        return k[: n.value], u[: n.value]

    # -- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(COMM_ID_BYTES)
        if not lib.scgpuCommGetUniqueId(buf):
            raise ScGpuError("scgpuCommGetUniqueId: " + lib.scgpuLastError(None).decode())
        return buf.raw

    def comm_init(self, n_ranks, rank, uid: bytes):
        buf = C.create_string_buffer(uid, COMM_ID_BYTES)
        self._ck(self.lib.scgpuCommInit(self.ctx, n_ranks, rank, buf), "scgpuCommInit")
        self.n_ranks = n_ranks
        self.rank = rank

    def enable_peer_gather(self, root=0, cap_entries=0):
        self._ck(self.lib.scgpuCommEnablePeerGather(self.ctx, root, cap_entries), "scgpuCommEnablePeerGather")

    def gather_visible(self, root=0):
        self._ck(self.lib.scgpuGatherVisible(self.ctx, root), "scgpuGatherVisible")

    def gathered_counts(self):
        out = np.zeros((self.n_ranks, self.n_views), np.uint32)
        self._ck(self.lib.scgpuGetGatheredCounts(self.ctx, _ptr(out), self.n_ranks), "scgpuGetGatheredCounts")
        return out

    def read_gathered_visible(self, view=0, out=None):
        n = C.c_uint32(0)
        self._ck(self.lib.scgpuReadGatheredVisible(self.ctx, view, None, 0, C.byref(n)), "scgpuReadGatheredVisible")
        if out is None:
            out = np.empty(n.value, np.uint32)
        self._ck(self.lib.scgpuReadGatheredVisible(self.ctx, view, _ptr(out), out.shape[0], C.byref(n)),
                 "scgpuReadGatheredVisible")
        return out[: n.value]
