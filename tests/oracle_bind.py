"""ctypes bindings of the parity checkers — TEST INFRASTRUCTURE ONLY.

  PortScene : oracle/libscoracle.so, our plain-C restatement (oracle/scoracle.c)
  RefScene  : oracle/_ref/libscref.so, the reference's own sources compiled headless (oracle/Makefile `ref`)

Both expose the same small interface as the GPU `scgpu.Scene`, so one scenario can be replayed on all three.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
PORT_LIB = ORACLE_DIR / "libscoracle.so"
REF_LIB = ORACLE_DIR / "_ref" / "libscref.so"
INVALID = 0xFFFFFFFF

_vp = C.c_void_p
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_port():
    if not PORT_LIB.exists() or PORT_LIB.stat().st_mtime < (ORACLE_DIR / "scoracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "port"], check=True, capture_output=True)
    return PORT_LIB


_port = None


def port_lib():
    global _port
    if _port is None:
        build_port()
        L = C.CDLL(str(PORT_LIB))
        L.sco_sinf.restype = C.c_float
        L.sco_sinf.argtypes = [C.c_float]
        L.sco_cosf.restype = C.c_float
        L.sco_cosf.argtypes = [C.c_float]
        L.sco_transform_system.restype = C.c_int64
        L.sco_transform_system.argtypes = [C.c_uint32, _vp, _vp, _vp, _vp, _vp]
        L.sco_culling_system.restype = None
        L.sco_culling_system.argtypes = [C.c_uint32, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _u32p, _vp, _u32p, _vp]
        L.sco_render_prep.restype = None
        L.sco_render_prep.argtypes = [C.c_uint32, _vp, _vp, _vp, _vp, C.c_uint32, _vp, _u32p, _u32p]
        L.sco_frame.restype = C.c_int64
        L.sco_frame.argtypes = [C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp]
        L.sco_sphere_in_frustum.restype = C.c_int
        L.sco_sphere_in_frustum.argtypes = [_vp, _vp, C.c_float]
        L.sco_mat4_perspective_rh_zo.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _vp]
        L.sco_mat4_rotation_xyz.argtypes = [C.c_float, C.c_float, C.c_float, _vp]
        L.sco_sincos_sweep.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.sco_expf.restype = C.c_float
        L.sco_expf.argtypes = [C.c_float]
        L.sco_atanf.restype = C.c_float
        L.sco_atanf.argtypes = [C.c_float]
        L.sco_atan2f.restype = C.c_float
        L.sco_atan2f.argtypes = [C.c_float, C.c_float]
        L.sco_unary_sweep.restype = C.c_uint64
        L.sco_unary_sweep.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, _vp]
        L.sco_atan2_sweep.restype = C.c_uint64
        L.sco_atan2_sweep.argtypes = [C.c_uint32, C.c_uint64, _vp]
        L.sco_lane_advance.restype = C.c_int
        L.sco_lane_advance.argtypes = [_vp, _u32p, C.POINTER(C.c_float), C.c_float, _vp, _vp]
        L.sco_lane_query_nearest.restype = C.c_uint32
        L.sco_lane_query_nearest.argtypes = [_vp, _vp, C.POINTER(C.c_float)]
        L.sco_traffic_ai_on_rails.restype = None
        L.sco_traffic_ai_on_rails.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_int, C.c_float,
                                              C.c_float, _vp]
        _port = L
    return _port


_ref = None


def ref_available() -> bool:
    return REF_LIB.exists()


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(str(REF_LIB))
        L.screfCreate.restype = _vp
        L.screfCreate.argtypes = [C.c_uint32]
        for name in ("screfDestroy", "screfRunTransform", "screfRunCulling"):
            getattr(L, name).argtypes = [_vp]
            getattr(L, name).restype = None
        L.screfCreateEntities.argtypes = [_vp, C.c_uint32, _vp]
        L.screfAddInstances.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp]
        L.screfDestroyEntities.argtypes = [_vp, C.c_uint32, _vp]
        L.screfSetLocal.argtypes = [_vp, C.c_uint32, _vp, _vp]
        L.screfSetParent.argtypes = [_vp, C.c_uint32, _vp, _vp]
        L.screfMarkDirty.argtypes = [_vp, C.c_uint32, _vp]
        L.screfSetViewProj.argtypes = [_vp, _vp]
        L.screfSetFreezeCulling.argtypes = [_vp, C.c_int]
        L.screfRunRenderPrep.argtypes = [_vp, C.c_uint32]
        L.screfAddCamera.restype = C.c_uint32
        L.screfAddCamera.argtypes = [_vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float]
        L.screfRunCamera.argtypes = [_vp, C.c_float]
        L.screfTransformCount.restype = C.c_uint32
        L.screfTransformCount.argtypes = [_vp]
        L.screfDenseEntities.restype = C.c_uint32
        L.screfDenseEntities.argtypes = [_vp, C.c_uint32, _vp]
        L.screfReadWorld.argtypes = [_vp, C.c_uint32, _vp, _vp]
        L.screfReadTransform.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp]
        L.screfGetViewProj.argtypes = [_vp, _vp]
        L.screfGetPlanes.argtypes = [_vp, _vp]
        L.screfGetCullStats.argtypes = [_vp, _u32p, _u32p, _u32p]
        for name in ("screfReadVisible", "screfReadCulled", "screfReadCandidates"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [_vp, C.c_uint32, _vp]
        L.screfGetRenderPrepStats.argtypes = [_vp, _u32p, _u32p]
        L.screfReadDraws.restype = C.c_uint32
        L.screfReadDraws.argtypes = [_vp, C.c_uint32, _vp]
        L.screfBuildDefaultScene.restype = C.c_uint32
        L.screfBuildDefaultScene.argtypes = [_vp, C.c_uint32]
        L.screfTimeFrame.argtypes = [_vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32, _vp, C.c_uint32, _f64p, _f64p, _f64p]
        L.screfMat4Perspective.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _vp]
        L.screfSphereInFrustum.restype = C.c_int
        L.screfSphereInFrustum.argtypes = [_vp, _vp, C.c_float]
        L.screfSinf.restype = C.c_float
        L.screfSinf.argtypes = [C.c_float]
        L.screfCosf.restype = C.c_float
        L.screfCosf.argtypes = [C.c_float]
        L.screfSinCosSweep.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.screfJobWorkers.restype = C.c_uint32
        L.screfLanesCreate.restype = _vp
        L.screfLanesCreate.argtypes = [C.c_float, C.c_float]
        L.screfLanesDestroy.argtypes = [_vp]
        L.screfLanesAddNode.restype = C.c_uint32
        L.screfLanesAddNode.argtypes = [_vp, _vp, _vp, C.c_float]
        L.screfLanesAddSegment.restype = C.c_uint32
        L.screfLanesAddSegment.argtypes = [_vp, C.c_uint32, C.c_uint32, _vp, C.c_int32, C.c_int32]
        L.screfLanesBuildSector.argtypes = [_vp, C.c_int32, C.c_int32, C.c_float]
        L.screfLanesRemoveSector.argtypes = [_vp, C.c_int32, C.c_int32]
        L.screfLanesSetActive.argtypes = [_vp, C.c_uint32, C.c_int]
        L.screfLanesCounts.argtypes = [_vp, _u32p, _u32p, _u32p]
        L.screfLanesExport.argtypes = [_vp] * 9 + [C.POINTER(C.c_float)]
        L.screfLaneAdvance.restype = C.c_int
        L.screfLaneAdvance.argtypes = [_vp, _u32p, C.POINTER(C.c_float), C.c_float, _vp, _vp]
        L.screfLaneQueryNearest.restype = C.c_uint32
        L.screfLaneQueryNearest.argtypes = [_vp, _vp, C.POINTER(C.c_float)]
        L.screfTrafficAddAgents.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]
        L.screfTrafficSetPlayer.argtypes = [_vp, C.c_uint32]
        L.screfRunTrafficAI.argtypes = [_vp, _vp, C.c_float, C.c_int, C.c_float, C.c_float]
        L.screfTrafficReadAgents.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]
        for name in ("screfExpf", "screfAtanf"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float]
        L.screfAtan2f.restype = C.c_float
        L.screfAtan2f.argtypes = [C.c_float, C.c_float]
        L.screfRendererSubmit.restype = C.c_uint32
        L.screfRendererSubmit.argtypes = [_vp, C.c_uint32, _vp, C.c_uint32, C.c_uint32, _vp, _vp]
        _ref = L
    return _ref


DRAW_ITEM_DTYPE = np.dtype(
    [("entity", "<u4"), ("meshId", "<u4"), ("materialId", "<u4"), ("_pad", "<u4"), ("model", "<f4", (16,))]
)

IDENTITY = np.eye(4, dtype=np.float32).ravel()


class PortScene:
    """SoA mirror of the Transform pool driven by the plain-C oracle. Pool bookkeeping (append, swap-remove) is
    done here in Python exactly as ComponentPool does it (sc_ecs.h:199-277)."""

    def __init__(self):
        self.L = port_lib()
        self.entity = np.zeros(0, np.uint32)
        self.parent = np.zeros(0, np.uint32)
        self.trs = np.zeros((0, 9), np.float32)
        self.world = np.zeros((0, 16), np.float32)
        self.dirty = np.zeros(0, np.uint8)
        self.flags = np.zeros(0, np.uint32)
        self.aabb = np.zeros((0, 6), np.float32)
        self.mesh_mat = np.zeros((0, 2), np.uint32)
        self.visible = []
        self.visible_slot = []
        self.culled = []
        self.recomputed = 0

    def _slot(self, handle):
        idx = np.nonzero(self.entity == np.uint32(handle))[0]
        return int(idx[0]) if idx.size else -1

    def spawn(self, entity, trs9, parent=None, aabb6=None, mesh_mat=None, flags=None):
        n = len(entity)
        self.entity = np.concatenate([self.entity, np.asarray(entity, np.uint32)])
        self.parent = np.concatenate([self.parent, np.full(n, INVALID, np.uint32) if parent is None else np.asarray(parent, np.uint32)])
        self.trs = np.concatenate([self.trs, np.asarray(trs9, np.float32).reshape(n, 9)])
        self.world = np.concatenate([self.world, np.tile(IDENTITY, (n, 1))])
        self.dirty = np.concatenate([self.dirty, np.ones(n, np.uint8)])
        self.flags = np.concatenate([self.flags, np.full(n, 3, np.uint32) if flags is None else np.asarray(flags, np.uint32)])
        unit = np.tile(np.array([-.5, -.5, -.5, .5, .5, .5], np.float32), (n, 1))
        self.aabb = np.concatenate([self.aabb, unit if aabb6 is None else np.asarray(aabb6, np.float32).reshape(n, 6)])
        self.mesh_mat = np.concatenate([self.mesh_mat, np.zeros((n, 2), np.uint32) if mesh_mat is None else np.asarray(mesh_mat, np.uint32).reshape(n, 2)])

    def despawn(self, entity):
        lookup = {int(e): i for i, e in enumerate(self.entity)}
        cnt = len(self.entity)
        arrays = [self.entity, self.parent, self.trs, self.world, self.dirty, self.flags, self.aabb, self.mesh_mat]
        for h in np.asarray(entity, np.uint32):
            s = lookup.get(int(h), -1)
            if s < 0:
                continue
            last = cnt - 1
            if s != last:
                moved = int(self.entity[last])
                for a in arrays:
                    a[s] = a[last]
                lookup[moved] = s
            del lookup[int(h)]
            cnt -= 1
        (self.entity, self.parent, self.trs, self.world, self.dirty, self.flags, self.aabb, self.mesh_mat) = [a[:cnt].copy() for a in arrays]

    def set_local(self, entity, trs9):
        t = np.asarray(trs9, np.float32).reshape(-1, 9)
        lookup = {int(e): i for i, e in enumerate(self.entity)}
        for h, row in zip(np.asarray(entity, np.uint32), t):
            s = lookup.get(int(h), -1)
            if s >= 0:
                self.trs[s] = row
                self.dirty[s] = 1

    def set_parent(self, entity, parent):
        lookup = {int(e): i for i, e in enumerate(self.entity)}
        for h, p in zip(np.asarray(entity, np.uint32), np.asarray(parent, np.uint32)):
            s = lookup.get(int(h), -1)
            if s >= 0:
                self.parent[s] = p
                self.dirty[s] = 1

    def mark_dirty(self, entity):
        lookup = {int(e): i for i, e in enumerate(self.entity)}
        for h in np.asarray(entity, np.uint32):
            s = lookup.get(int(h), -1)
            if s >= 0:
                self.dirty[s] = 1

    def mark_all_dirty(self):
        self.dirty[:] = 1

    def update(self, view_projs, freeze=False, skip_transform=False):
        n = len(self.entity)
        for name in ("entity", "parent", "trs", "world", "dirty", "flags", "aabb", "mesh_mat"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name)))
        if not skip_transform:
            self.recomputed = self.L.sco_transform_system(n, _p(self.entity), _p(self.parent), _p(self.trs), _p(self.world), _p(self.dirty))
        vps = np.ascontiguousarray(view_projs, np.float32).reshape(-1, 16)
        self.visible, self.visible_slot, self.culled = [], [], []
        for v in range(vps.shape[0]):
            vis = np.zeros(n, np.uint32)
            slot = np.zeros(n, np.uint32)
            cul = np.zeros(n, np.uint32)
            nv = C.c_uint32(0)
            nc = C.c_uint32(0)
            self.L.sco_culling_system(n, _p(self.entity), _p(self.flags), _p(self.world), _p(self.aabb), _p(vps[v]), 1 if freeze else 0,
                                      _p(vis), C.byref(nv), _p(cul), C.byref(nc), _p(slot))
            self.visible.append(vis[: nv.value].copy())
            self.visible_slot.append(slot[: nv.value].copy())
            self.culled.append(cul[: nc.value].copy())

    def read_world(self, entity):
        lookup = {int(e): i for i, e in enumerate(self.entity)}
        return np.stack([self.world[lookup[int(h)]] for h in np.asarray(entity, np.uint32)]) if len(entity) else np.zeros((0, 16), np.float32)

    def read_draw_items(self, view=0, max_draws=0):
        slots = np.ascontiguousarray(self.visible_slot[view])
        out = np.zeros(len(slots), DRAW_ITEM_DTYPE)
        e = C.c_uint32(0)
        d = C.c_uint32(0)
        self.L.sco_render_prep(len(slots), _p(slots), _p(self.entity), _p(np.ascontiguousarray(self.mesh_mat)), _p(self.world), max_draws, _p(out),
                               C.byref(e), C.byref(d))
        return out[: e.value], e.value, d.value


class RefScene:
    """The reference's own World + systems (oracle/_ref/libscref.so)."""

    def __init__(self, workers=0):
        self.L = ref_lib()
        self.w = self.L.screfCreate(workers)
        assert self.w
        self.visible, self.culled = [], []

    def close(self):
        if self.w:
            self.L.screfDestroy(self.w)
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def create_entities(self, n):
        out = np.zeros(n, np.uint32)
        self.L.screfCreateEntities(self.w, n, _p(out))
        return out

    def spawn(self, entity, trs9, parent=None, aabb6=None, mesh_mat=None, flags=None):
        e = np.ascontiguousarray(entity, np.uint32)
        n = len(e)
        t = np.ascontiguousarray(trs9, np.float32).reshape(n, 9)
        p = None if parent is None else np.ascontiguousarray(parent, np.uint32)
        b = np.tile(np.array([-.5, -.5, -.5, .5, .5, .5], np.float32), (n, 1)) if aabb6 is None else np.ascontiguousarray(aabb6, np.float32).reshape(n, 6)
        m = None if mesh_mat is None else np.ascontiguousarray(mesh_mat, np.uint32).reshape(n, 2)
        f = None if flags is None else np.ascontiguousarray(flags, np.uint32)
        self.L.screfAddInstances(self.w, n, _p(e), _p(p), _p(t), _p(b), _p(m), _p(f))

    def despawn(self, entity):
        e = np.ascontiguousarray(entity, np.uint32)
        self.L.screfDestroyEntities(self.w, len(e), _p(e))

    def set_local(self, entity, trs9):
        e = np.ascontiguousarray(entity, np.uint32)
        t = np.ascontiguousarray(trs9, np.float32).reshape(len(e), 9)
        self.L.screfSetLocal(self.w, len(e), _p(e), _p(t))

    def set_parent(self, entity, parent):
        e = np.ascontiguousarray(entity, np.uint32)
        p = np.ascontiguousarray(parent, np.uint32)
        self.L.screfSetParent(self.w, len(e), _p(e), _p(p))

    def mark_dirty(self, entity):
        e = np.ascontiguousarray(entity, np.uint32)
        self.L.screfMarkDirty(self.w, len(e), _p(e))

    def mark_all_dirty(self):
        self.mark_dirty(self.dense_entities())

    def dense_entities(self):
        n = self.L.screfTransformCount(self.w)
        out = np.zeros(n, np.uint32)
        self.L.screfDenseEntities(self.w, n, _p(out))
        return out

    def update(self, view_projs, freeze=False, skip_transform=False):
        if not skip_transform:
            self.L.screfRunTransform(self.w)
        self.L.screfSetFreezeCulling(self.w, 1 if freeze else 0)
        vps = np.ascontiguousarray(view_projs, np.float32).reshape(-1, 16)
        n = self.L.screfTransformCount(self.w)
        self.visible, self.culled = [], []
        for v in range(vps.shape[0]):
            self.L.screfSetViewProj(self.w, _p(vps[v]))
            self.L.screfRunCulling(self.w)
            vis = np.zeros(n, np.uint32)
            cul = np.zeros(n, np.uint32)
            nv = self.L.screfReadVisible(self.w, n, _p(vis))
            nc = self.L.screfReadCulled(self.w, n, _p(cul))
            self.visible.append(vis[:nv].copy())
            self.culled.append(cul[:nc].copy())
        # leave view 0's result in CullingState for render prep
        if vps.shape[0] > 1:
            self.L.screfSetViewProj(self.w, _p(vps[0]))
            self.L.screfRunCulling(self.w)

    def read_world(self, entity):
        e = np.ascontiguousarray(entity, np.uint32)
        out = np.zeros((len(e), 16), np.float32)
        self.L.screfReadWorld(self.w, len(e), _p(e), _p(out))
        return out

    def read_parents(self, entity):
        e = np.ascontiguousarray(entity, np.uint32)
        out = np.zeros(len(e), np.uint32)
        self.L.screfReadTransform(self.w, len(e), _p(e), _p(out), None, None)
        return out

    def read_draw_items(self, view=0, max_draws=0):
        assert view == 0
        self.L.screfRunRenderPrep(self.w, max_draws)
        e = C.c_uint32(0)
        d = C.c_uint32(0)
        self.L.screfGetRenderPrepStats(self.w, C.byref(e), C.byref(d))
        out = np.zeros(e.value, DRAW_ITEM_DTYPE)
        self.L.screfReadDraws(self.w, e.value, _p(out))
        return out, e.value, d.value

    def planes(self):
        out = np.zeros(24, np.float32)
        self.L.screfGetPlanes(self.w, _p(out))
        return out.reshape(6, 4)


# ---- the renderer's per-frame CPU sort of the draws, restated (test infrastructure) ------------

def renderer_sorted_draws(draws, material_pipeline, mesh_count):
    """What src/engine/src/sc_vk.cpp:1843-1905 does with RenderFrameData::draws, in numpy:
       :1847-1850  skip draws with meshId >= m_meshes.size() or without a material,
       :1854-1864  sort by (material->pipelineId, materialId, meshId) — std::sort there (unstable); stable here, which
                   is one of the orders std::sort may produce,
       :1866-1905  bind pipeline / material / mesh on change: the maximal runs of equal (pipeline, material, mesh).
    draws: structured array with meshId / materialId (scgpu.DRAW_ITEM_DTYPE layout or the oracle's dict form).
    Returns (order: indices into draws, runs: list of (pipeline, material, mesh, first, count))."""
    mesh = np.asarray(draws["meshId"], np.int64)
    mat = np.asarray(draws["materialId"], np.int64)
    mp = np.asarray(material_pipeline, np.int64)
    known = mat < len(mp)
    pipe = np.where(known, mp[np.minimum(mat, max(len(mp) - 1, 0))] if len(mp) else 0xFFFFFFFF, 0xFFFFFFFF)
    keep = (mesh < mesh_count) & known & (pipe != 0xFFFFFFFF)
    idx = np.nonzero(keep)[0]
    order = idx[np.lexsort((mesh[idx], mat[idx], pipe[idx]))]  # lexsort: last key is primary; stable
    runs = []
    for pos, i in enumerate(order):
        k = (int(pipe[i]), int(mat[i]), int(mesh[i]))
        if runs and tuple(runs[-1][:3]) == k:
            runs[-1][4] += 1
        else:
            runs.append([k[0], k[1], k[2], pos, 1])
    return order, [tuple(r) for r in runs]


def ref_renderer_submit(draws, material_pipeline, mesh_count):
    """The reference's OWN draw submission block (src/engine/src/sc_vk.cpp:1841-1912, compiled into oracle/_ref by
    ref_shim/scref_renderer.cpp) run over `draws` (DRAW_ITEM_DTYPE records). Returns (order, binds): order[k] = index of
    the k-th vkCmdDrawIndexed'ed item, binds[k] = what the loop bound right before it (1 pipeline | 2 material | 4 mesh)."""
    d = np.ascontiguousarray(draws, DRAW_ITEM_DTYPE)
    mp = np.ascontiguousarray(material_pipeline, np.uint32)
    order = np.zeros(len(d), np.uint32)
    binds = np.zeros(len(d), np.uint8)
    n = ref_lib().screfRendererSubmit(_p(d), len(d), _p(mp), len(mp), int(mesh_count), _p(order), _p(binds))
    return order[:n], binds[:n]


def check_against_renderer(draws, material_pipeline, mesh_count, order, runs, what=""):
    """(order, runs) — ours, or the numpy restatement's — against what the reference's loop submitted: the same set of
    draws, the same (pipeline, material, mesh) key at every position, a run starting exactly where the loop bound
    something. std::sort is unstable: WHICH of several equal-key items stands at a position is not defined by the
    reference, so inside a run only the multiset of items is compared."""
    ro, rb = ref_renderer_submit(draws, material_pipeline, mesh_count)
    assert len(ro) == len(order), (what, len(ro), len(order))
    mesh, mat = np.asarray(draws["meshId"], np.int64), np.asarray(draws["materialId"], np.int64)
    mp = np.asarray(material_pipeline, np.int64)
    key = lambda idx: (mp[mat[idx]] << 58) | (mat[idx] << 29) | mesh[idx]
    assert np.array_equal(key(ro), key(np.asarray(order))), what + ": key sequence differs from the reference's submission order"
    starts = np.flatnonzero(rb != 0)
    assert [int(r[3]) for r in runs] == [int(x) for x in starts], what + ": run starts != the loop's bind points"
    assert np.array_equal(np.sort(ro), np.sort(np.asarray(order))), what + ": kept set differs"
    for r in runs:
        a, b = int(r[3]), int(r[3]) + int(r[4])
        assert np.array_equal(np.sort(ro[a:b]), np.sort(np.asarray(order)[a:b])), what + ": items of a run differ"


# ---- traffic on rails -----------------------------------------------------------------------

class ScoLaneGraph(C.Structure):
    _fields_ = [("nNodes", C.c_uint32), ("nSegments", C.c_uint32), ("nodePos", _vp), ("nodeSpeedLimit", _vp),
                ("nodeConnOffset", _vp), ("nodeConn", _vp), ("segNodes", _vp), ("segDir", _vp), ("segLength", _vp),
                ("segActive", _vp), ("defaultSpeedLimit", C.c_float)]


LANE_KEYS = ("node_pos", "node_speed", "conn_offset", "conn", "seg_nodes", "seg_dir", "seg_len", "seg_active")


def sco_graph(g):
    """ScoLaneGraph over the arrays of a lane-graph dict (keys LANE_KEYS + default_speed); the dict must outlive it."""
    return ScoLaneGraph(len(g["node_speed"]), len(g["seg_len"]), _p(g["node_pos"]), _p(g["node_speed"]), _p(g["conn_offset"]),
                        _p(g["conn"]), _p(g["seg_nodes"]), _p(g["seg_dir"]), _p(g["seg_len"]), _p(g["seg_active"]),
                        float(g["default_speed"]))


def port_traffic_step(g, agents, trs9, dt, brake=None, skip=None, debug=None):
    """sco_traffic_ai_on_rails in place on agents = dict(lane u32, s f32, speed f32, look f32) and trs9 [n,9].
    debug = (lookAheadDist, speedMultiplier) or None. Returns the moved mask."""
    L = port_lib()
    n = len(agents["lane"])
    moved = np.zeros(n, np.uint8)
    sg = sco_graph(g)
    L.sco_traffic_ai_on_rails(C.byref(sg), n, _p(agents["lane"]), _p(agents["s"]), _p(agents["speed"]), _p(agents["look"]),
                              _p(trs9), _p(brake), _p(skip), dt, 1 if debug else 0, debug[0] if debug else 0.0,
                              debug[1] if debug else 0.0, _p(moved))
    return moved


class RefLanes:
    """The reference's TrafficLaneGraph (oracle/_ref), built through its own addNode / addSegment /
    buildProceduralForSector, exported as the flat arrays every implementation consumes."""

    def __init__(self, lane_width=3.5, speed_limit=12.0):
        self.L = ref_lib()
        self.h = self.L.screfLanesCreate(lane_width, speed_limit)

    def close(self):
        if self.h:
            self.L.screfLanesDestroy(self.h)
            self.h = None

    def add_node(self, pos, direction, speed):
        return self.L.screfLanesAddNode(self.h, _p(np.asarray(pos, np.float32)), _p(np.asarray(direction, np.float32)), speed)

    def add_segment(self, a, b, direction, owner=(0, 0)):
        return self.L.screfLanesAddSegment(self.h, a, b, _p(np.asarray(direction, np.float32)), owner[0], owner[1])

    def build_sector(self, x, z, size=64.0):
        self.L.screfLanesBuildSector(self.h, x, z, size)

    def remove_sector(self, x, z):
        self.L.screfLanesRemoveSector(self.h, x, z)

    def set_active(self, seg, active):
        self.L.screfLanesSetActive(self.h, seg, 1 if active else 0)

    def export(self):
        nn, ns, nc = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        self.L.screfLanesCounts(self.h, C.byref(nn), C.byref(ns), C.byref(nc))
        g = dict(node_pos=np.zeros((nn.value, 3), np.float32), node_speed=np.zeros(nn.value, np.float32),
                 conn_offset=np.zeros(nn.value + 1, np.uint32), conn=np.zeros(max(nc.value, 1), np.uint32)[: nc.value],
                 seg_nodes=np.zeros((ns.value, 2), np.uint32), seg_dir=np.zeros((ns.value, 3), np.float32),
                 seg_len=np.zeros(ns.value, np.float32), seg_active=np.zeros(ns.value, np.uint8))
        d = C.c_float(0)
        self.L.screfLanesExport(self.h, *[_p(g[k]) for k in LANE_KEYS], C.byref(d))
        g["default_speed"] = np.float32(d.value)
        return g


def ref_traffic_frames(lanes, agents, trs9, dts, debug=None):
    """Runs the reference's own TrafficAISystem (physics == nullptr, streaming == nullptr) over a World holding one
    PlayerVehicle and the given on-rails agents for every dt in dts. Returns per frame (lane, s, speed, look, trs9)."""
    R = RefScene(1)
    L = R.L
    n = len(agents["lane"])
    e = R.create_entities(n + 1)
    player, ag = e[:1], np.ascontiguousarray(e[1:])
    R.spawn(player, np.array([[0, 0, 0, 0, 0, 0, 1, 1, 1]], np.float32))
    R.spawn(ag, trs9)
    L.screfTrafficSetPlayer(R.w, int(player[0]))
    L.screfTrafficAddAgents(R.w, n, _p(ag), _p(agents["lane"]), _p(agents["s"]), _p(agents["speed"]), _p(agents["look"]))
    out = []
    for dt in dts:
        L.screfRunTransform(R.w)  # clears Transform::dirty, so that after the AI pass dirty == "moved this frame"
        L.screfRunTrafficAI(R.w, lanes.h, dt, 1 if debug else 0, debug[0] if debug else 0.0, debug[1] if debug else 0.0)
        lane, s, v, look = np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        L.screfTrafficReadAgents(R.w, n, _p(ag), _p(lane), _p(s), _p(v), _p(look))
        t = np.zeros((n, 9), np.float32)
        dirty = np.zeros(n, np.uint8)
        L.screfReadTransform(R.w, n, _p(ag), None, _p(t), _p(dirty))
        out.append((lane, s, v, look, t, dirty))
    R.close()
    return out
