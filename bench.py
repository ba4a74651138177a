#!/usr/bin/env python
"""bench.py — scene-update hot path benchmark (transform -> cull (V views) -> compact), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (libscgpu.so through the C ABI)
  python bench.py --impl reference ...                          the reference's own CPU code (oracle/_ref), rank 0 only

A "step" is one frame over the whole per-GPU instance set with EVERY instance dirty (worst case: each one is
re-transformed, culled against all views and compacted). Workload at any N: the north-star per-GPU share —
16 Mi instances per GPU in depth-4 groups, main camera + 4 shadow cascades (BASELINE.json configs[2]; 8 GPUs of
it = configs[3], 128 Mi instances sharded by world cell). Weak scaling: every rank owns a contiguous block of
world cells; the only exchange is the gather of the compacted lists to rank 0 (inside the timed step for N>1).

Beside the headline numbers the same line carries
  "clean_frame"    the same scene with nothing dirty (static city, the camera moves): the cull-only kernel, 96 B per instance;
  "partial_dirty"  the same scene with 30 % of the instances dirty per frame (TransformSystem only recomputes
                   t.dirty || parentDirty, sc_ecs.cpp:178-210): device time and the mixed 132 / 96-byte roofline;
  "churn"          BASELINE.json configs[4], the per-GPU share: 8 Mi instances, per frame 10 % despawn + 10 % spawn +
                   30 % setLocal through the C ABI from host buffers, >= 100 timed frames, the fused kernel's time series
                   (it must stay flat: the device layout no longer follows the pool's swap-with-last order);
  "gather"         (N > 1) the gathered per-view totals on rank 0, checked against the sum of the ranks' own counts and
                   the concatenation of their lists.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))

METRIC = "instances transformed+culled/sec"
UNIT = "instances/s"
DEFAULT_INSTANCES = 16 * 1024 * 1024 - 4096  # per GPU; < 2^24 entity indices per World shard (sc_ecs.h:18-20)
ALG_BYTES_DIRTY = 132   # read TRS 36 + parent 4 + flags 4 + AABB 24, write world 64
ALG_BYTES_CLEAN = 96    # read flags 4 + parent 4 + world 64 + AABB 24, no write
REF_SAMPLE = 2_000_000  # instances per step of the bounded cpu_baseline sample (N = 1 line of our arm)
REF_WORLD_MAX = 2_000_000  # The reference's job-payload arena holds ~11.18 M culling jobs' worth per frame (BASELINE.md §2:
                           # beyond it CullingSystem never returns) and the baseline culls 5 views per frame, so a World
                           # may hold at most ~2.2 M instances here. Larger scenes run as consecutive Worlds of at most
                           # this many instances, times summed.
CHURN_INSTANCES = 8 * 1024 * 1024
TIMING_EVERY = int(__import__("os").environ.get("SCGPU_BENCH_TIMING_EVERY", "4"))  # steps between two steps with per-kernel events


def block_shift(rank, world, side):
    """x offset of rank r's block of world cells: the blocks stand side by side along x and the camera sits over the middle
    of the whole world, i.e. (for N > 1) on the seam between two ranks' blocks — the visible lists come from more than one
    GPU and none of them need be the submitting rank's. N = 1: no shift."""
    from scgpu import scenes
    return np.float32((rank - (world - 1) / 2.0) * side * scenes.SECTOR_SIZE)


def build_scene(n, rank, seed=424242, world=1):
    from scgpu import scenes
    if os.environ.get("SCGPU_BENCH_WORKLOAD", "hier") == "flat":  # diagnostic only; the reported workload is "hier"
        sc = scenes.city_flat(n, seed=seed + 7919 * rank)
        return sc
    sc = scenes.city_hier(n, seed=seed + 7919 * rank)
    roots = sc["parent"] < 0
    sc["trs9"][roots, 0] += block_shift(rank, world, sc["side"])
    return sc


def views_for(sc, n_views):
    from scgpu import scenes
    # camera over the origin sector of rank 0's block, like the sandbox's (32, 6, 44) yaw pi (main.cpp:82-88)
    return scenes.standard_views(n_views)


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons during the timed region (nvidia-smi's clocks line, via NVML)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.armed = False  # NVML is initialised and the thread started BEFORE the barrier; sampling starts when armed
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._halt.is_set():
            if not self.armed:
                time.sleep(0.0005)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)  # the timed region lasts ~10 ms: sample as fast as NVML answers

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None"""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline worker (runs in a subprocess with the libm variant pinned)
# --------------------------------------------------------------------------------------------------------

def _world_cuts(sc, n, limit):
    """[a, b) ranges of at most `limit` instances cut at group boundaries (a group never straddles two Worlds)"""
    roots = np.nonzero(sc["parent"] < 0)[0]
    cuts, a = [], 0
    while a < n:
        b = min(n, a + limit)
        if b < n:
            b = int(roots[np.searchsorted(roots, b, side="right") - 1])
        cuts.append((a, b))
        a = b
    return cuts


def cpu_worker(args):
    """Times the reference's own TransformSystem -> CullingSystem x V -> RenderPrepStreamingSystem over `--sample-instances`
    instances of the bench workload, as consecutive Worlds of at most REF_WORLD_MAX instances whose times are summed.
    Prints one JSON object."""
    sys.path.insert(0, str(ROOT / "tests"))
    import ctypes as C
    import oracle_bind
    n = args.sample_instances
    sc = build_scene(n, 0)
    vps = views_for(sc, args.views)
    from scgpu import scenes
    out = {"cores": os.cpu_count()}
    if oracle_bind.ref_available():
        cuts = _world_cuts(sc, n, REF_WORLD_MAX)
        worlds = []
        for a, b in cuts:
            r = oracle_bind.RefScene(0)
            e = r.create_entities(b - a)
            lp = np.where(sc["parent"][a:b] >= 0, sc["parent"][a:b] - a, -1)
            r.spawn(e, sc["trs9"][a:b], scenes.parent_handles(lp, e), sc["aabb6"][a:b], sc["mesh_mat"][a:b], sc["flags"][a:b])
            worlds.append((r, np.ascontiguousarray(e)))
        L = worlds[0][0].L
        vp = np.ascontiguousarray(vps, np.float32)
        tT, tC, tP = C.c_double(), C.c_double(), C.c_double()
        per_step = []
        for it in range(args.warmup + args.steps):
            acc = np.zeros(3)
            for r, ee in worlds:
                L.screfTimeFrame(r.w, 1, args.views, vp.ctypes.data_as(C.c_void_p), len(ee), ee.ctypes.data_as(C.c_void_p), 0,
                                 C.byref(tT), C.byref(tC), C.byref(tP))
                acc += (tT.value, tC.value, tP.value)
            if it >= args.warmup:
                per_step.append(acc)
        a = np.array(per_step)
        out.update(kind="reference", threads=int(L.screfJobWorkers()) + 1,
                   transform_ms=float(a[:, 0].mean() * 1e3), cull_ms=float(a[:, 1].mean() * 1e3),
                   prep_ms=float(a[:, 2].mean() * 1e3), worlds=[int(b - a_) for a_, b in cuts])
        sec = float(a.sum(axis=1).mean())
        for r, _ in worlds:
            r.close()
        out["sample"] = (f"{n} instances of the same depth-4 city, {args.views} views, all dirty, per step; "
                         f"{len(cuts)} consecutive World(s) of <= {REF_WORLD_MAX} instances, times summed")
    else:
        p = oracle_bind.PortScene()
        e = np.arange(n, dtype=np.uint32)
        p.spawn(e, sc["trs9"], scenes.parent_handles(sc["parent"], e), sc["aabb6"], sc["mesh_mat"], sc["flags"])
        times = []
        for it in range(args.warmup + args.steps):
            p.mark_all_dirty()
            t0 = time.perf_counter()
            p.update(vps)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
        out.update(kind="port", threads=1)
        sec = float(np.mean(times))
        out["sample"] = f"{n} instances of the same depth-4 city, {args.views} views, all dirty, per step (plain-C oracle, 1 thread)"
    out.update(value=n / sec, unit=UNIT, ms_per_step=sec * 1e3, sample_instances=n)
    print(json.dumps(out), flush=True)


def run_cpu_worker(sample, views, steps, warmup, timeout=900):
    env = dict(os.environ)
    env["GLIBC_TUNABLES"] = "glibc.cpu.hwcaps=-FMA,-AVX2"  # pins libm's sinf/cosf variant
    cmd = [sys.executable, str(ROOT / "bench.py"), "--cpu-worker", "--sample-instances", str(sample), "--views",
           str(views), "--steps", str(steps), "--warmup", str(warmup)]
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            return {"error": (r.stderr or r.stdout)[-400:]}
        return json.loads(line[-1])
    except Exception as ex:  # noqa
        return {"error": repr(ex)}


def reference_arm(args, rank, world):
    """The reference's own CPU implementation on OUR arm's config: the whole per-GPU instance set (16.77 M), as
    consecutive Worlds of <= 2 M (5 views per frame). --quick falls back to the bounded 2 M sample."""
    if rank != 0:
        return
    sample = args.ref_sample if args.quick else args.instances
    steps = args.steps if args.quick else min(args.steps, 6)   # ~4 s per 16.77 M step: keep the run within minutes
    warm = min(args.warmup, 1)
    res = run_cpu_worker(sample, args.views, steps, warm, timeout=3000)
    if "error" in res:
        print(json.dumps({"impl": "reference", "unavailable": res["error"].replace("\n", " ")[:200]}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, res["sample_instances"], sample=True),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res.get("threads", res["cores"]),
                         "kind": res["kind"], "sample": res["sample"],
                         "stages_ms": {k: res[k] for k in ("transform_ms", "cull_ms", "prep_ms") if k in res},
                         "host_cpus": res["cores"], "worlds": res.get("worlds")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_per_gpu, sample=False):
    return {
        "workload": "synthetic city, depth-4 groups (vehicles+wheels, peds+attachments), main view + 4 shadow "
                    "cascades, all instances dirty every step (BASELINE.json configs[2]; x8 GPUs = configs[3])",
        "instances_per_gpu": int(n_per_gpu), "views": args.views, "dirty_fraction": 1.0,
        "sharding": ("world cell blocks side by side, one process per GPU, the camera over the seam in the middle of the "
                     "world; visible lists gathered to rank 0 through "
                     + ("NVLink peer memory" if os.environ.get("SCGPU_GATHER", "peer") == "peer" else "NCCL send/recv"))
        if args.gpus > 1 else "single GPU",
        "l2": "inputs (>= 2 GB per step) exceed the 126 MB L2; no explicit flush" if not sample else "cpu run",
    }


# --------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------

class Dist:
    """torch.distributed when WORLD_SIZE > 1, no-ops otherwise"""

    def __init__(self, world, local_rank):
        self.world = world
        self.dist = None
        if world > 1:
            import torch
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        import torch
        if self.dist:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max(self, x):
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_objects(self, obj):
        if not self.dist:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def comm_setup(scene, D, rank):
    import scgpu
    if D.world > 1:
        uid = [scgpu.Scene.comm_unique_id() if rank == 0 else None]
        D.dist.broadcast_object_list(uid, src=0)
        scene.comm_init(D.world, rank, uid[0])
        if os.environ.get("SCGPU_GATHER", "peer") == "peer":  # lists travel through NVLink peer memory, not NCCL
            scene.enable_peer_gather(0)


def check_gather(scene, D, rank, views):
    """rank 0 reads what the gather delivered and compares it with what the ranks hold themselves"""
    if D.world == 1:
        return None
    scene.update(0)
    scene.gather_visible(0)
    local_counts = [int(scene.counts().visible[v]) for v in range(views)]
    local_lists = [scene.read_visible(v) for v in range(views)]
    scene.synchronize()
    import zlib
    mine = {"counts": local_counts, "crc": [zlib.crc32(l.tobytes()) for l in local_lists]}
    everyone = D.gather_objects(mine)
    out = None
    if rank == 0:
        g = scene.gathered_counts()
        ok = True
        totals = []
        for v in range(views):
            per_rank = [everyone[r]["counts"][v] for r in range(D.world)]
            ok &= [int(x) for x in g[:, v]] == per_rank
            lst = scene.read_gathered_visible(v)
            ok &= len(lst) == sum(per_rank)
            off = 0
            for r in range(D.world):   # concatenated in rank order: every rank's slice must be that rank's own list
                ok &= zlib.crc32(lst[off:off + per_rank[r]].tobytes()) == everyone[r]["crc"][v]
                off += per_rank[r]
            totals.append(int(sum(per_rank)))
        out = {"gather_checked": bool(ok), "gathered_visible_per_view": totals,
               "rank0_local_visible_per_view": local_counts,
               "check": "gathered counts == every rank's own counts; gathered lists == rank-ordered concatenation of the ranks' lists (crc32 per slice)"}
        if not ok:
            raise SystemExit("bench.py: the gathered visible lists differ from the ranks' own lists")
    D.barrier()
    return out


def partial_dirty_leg(scene, sc, n, args, D, stream, torch):
    """30 % of the instances get a new local TRS per frame from a DEVICE-resident producer (scgpuSetLocalDevice), then the
    update: what TransformSystem's dirty test (sc_ecs.cpp:178-210) buys. Returns the dict for the JSON line."""
    rng = np.random.default_rng(12345)
    m = (3 * n) // 10
    idx = np.sort(rng.choice(n, m, replace=False)).astype(np.uint32)
    trs = sc["trs9"][idx].copy()
    trs[:, 0] += np.float32(0.125)
    dev = torch.device("cuda", torch.cuda.current_device())
    d_e = torch.from_numpy(idx.view(np.int32)).to(dev)
    d_t = torch.from_numpy(trs).to(dev)
    torch.cuda.synchronize()
    steps = max(5, min(args.steps, 20))

    def step():
        scene.set_local_device(m, d_e.data_ptr(), d_t.data_ptr())
        scene.update(0)
        if D.world > 1:
            scene.gather_visible(0)

    for _ in range(3):
        step()
    D.barrier()
    scene.enable_timings(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    D.barrier()
    ms = D.max(ev0.elapsed_time(ev1) / steps)
    k_ms, u_ms = scene.read_timings(steps)
    rec = int(scene.counts().recomputed)
    peak, _ = measured_peak()
    alg = ALG_BYTES_DIRTY * rec + ALG_BYTES_CLEAN * (n - rec)
    k = float(np.mean(k_ms))
    return {"dirty_fraction_set": 0.3, "recomputed_per_frame": rec, "recomputed_fraction": rec / n, "steps": steps,
            "ms_per_step": ms, "value": n * D.world / (ms * 1e-3), "unit": UNIT, "kernel_ms_avg": k,
            "update_ms_avg": float(np.mean(u_ms)),
            "roofline": {"bound": "hbm", "achieved": alg / (k * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (k * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(alg),
                         "formula": "132 B x recomputed + 96 B x clean (SURVEY.md 8d)"},
            "note": "dirty instances chosen at random (every hierarchy window is partly dirty, children inherit); the step "
                    "includes the device-side setLocal of the 30 % (k_set_local), kernel_ms is the fused kernel alone"}


def clean_leg(scene, n, args, D, stream, torch):
    """Frames in which nothing is dirty (a static city under a moving camera): the library sees that no delta call came
    since the last transforming update and runs the cull-only kernel: 96 B per instance (AABB + flags + stored matrix)."""
    steps = max(5, min(args.steps, 20))

    def step():
        scene.update(0)
        if D.world > 1:
            scene.gather_visible(0)

    for _ in range(3):
        step()
    D.barrier()
    scene.enable_timings(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    D.barrier()
    ms = D.max(ev0.elapsed_time(ev1) / steps)
    k_ms, u_ms = scene.read_timings(steps)
    c = scene.counts()
    assert int(c.recomputed) == 0
    peak, _ = measured_peak()
    k = float(np.mean(k_ms))
    alg = ALG_BYTES_CLEAN * n
    return {"dirty_fraction": 0.0, "steps": steps, "ms_per_step": ms, "value": n * D.world / (ms * 1e-3), "unit": UNIT,
            "kernel": "k_cull_only<V>", "kernel_ms_avg": k, "update_ms_avg": float(np.mean(u_ms)),
            "visible_per_view": [int(c.visible[v]) for v in range(args.views)],
            "roofline": {"bound": "hbm", "achieved": alg / (k * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (k * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(alg),
                         "formula": "96 B x instances (SURVEY.md 8d: flags 4 + parent 4 + world 64 + AABB 24)"}}


def churn_leg(args, D, rank, local_rank, torch):
    """BASELINE.json configs[4], per-GPU share: 8 Mi instances in depth-4 groups, 5 views; per frame ~10 % of the instances
    despawn as whole groups, as many spawn (same entity indices, next generation, like the reference's LIFO index
    reuse), 30 % of the rest get a new local TRS — all through the C ABI from pinned HOST buffers. >= 100 timed frames."""
    import scgpu
    from scgpu import scenes
    n = args.churn_instances
    frames, warm = args.churn_frames, 4
    cohorts = int(os.environ.get("SCGPU_CHURN_COHORTS", "10"))  # 1 / cohorts of the sectors stream out and in per frame
    sc = scenes.city_hier(n, seed=99 + 7919 * rank)
    sc["trs9"][sc["parent"] < 0, 0] += block_shift(rank, D.world, sc["side"])
    rng = np.random.default_rng(5 + rank)
    # The unit of streaming is the world SECTOR (64 m cell), as in the reference's WorldPartition (pumpUnloadQueue destroys a
    # sector's entities, pumpCompletedLoads creates one's, sc_world_partition.cpp:839-1034): every sector belongs to one of
    # ten cohorts, at random; frame f unloads cohort f % 10 and loads it again — the same content, but the sectors arrive
    # in a fresh random order, so a new sector lands in whatever holes an old one left, not in its own.
    sec_key = sc["sector"][:, 0].astype(np.int64) * (1 << 20) + sc["sector"][:, 1].astype(np.int64)
    sec_start = np.concatenate([[0], np.nonzero(np.diff(sec_key))[0] + 1])   # the scene is laid out sector by sector
    sec_len = np.diff(np.append(sec_start, n))
    sco = rng.integers(0, cohorts, size=len(sec_start))
    cohort_of = np.repeat(sco, sec_len).astype(np.uint8)
    index = np.arange(n, dtype=np.uint32)                     # entity index == initial position, reused by every generation
    gen = np.zeros(cohorts, np.uint32)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    tmpl = []
    for c in range(cohorts):
        secs = np.nonzero(sco == c)[0]
        tmpl.append(dict(secs=secs, ix=np.nonzero(cohort_of == c)[0].astype(np.uint32)))
    moved_pat = [np.sort(rng.choice(n, (3 * n) // 10, replace=False)).astype(np.uint32) for _ in range(4)]
    moved_trs = [pin(sc["trs9"][p] + np.float32(0.25)) for p in moved_pat]

    def spawn_batch(c):
        """cohort c's sectors in a fresh random order: handles, parents, components (host prep, outside the timed calls)"""
        secs = rng.permutation(tmpl[c]["secs"])
        ix = np.concatenate([np.arange(sec_start[k], sec_start[k] + sec_len[k], dtype=np.int64) for k in secs])
        pos = np.full(n, -1, np.int64)
        pos[ix] = np.arange(len(ix))
        lp = np.where(sc["parent"][ix] >= 0, pos[np.maximum(sc["parent"][ix], 0)], -1)
        fresh = pin(ix.astype(np.uint32) | (gen[c] << np.uint32(24)))
        return (fresh, pin(scenes.parent_handles(lp, fresh)), pin(sc["trs9"][ix]), pin(sc["aabb6"][ix]), pin(sc["mesh_mat"][ix]),
                pin(sc["flags"][ix]))

    scene = scgpu.Scene(n + n // 8, max_views=args.views, device=local_rank, max_entity_index=n)
    scene.spawn(index, sc["trs9"], scenes.parent_handles(sc["parent"], index), sc["aabb6"], sc["mesh_mat"], sc["flags"])
    scene.set_views(scenes.standard_views(args.views))
    comm_setup(scene, D, rank)
    scene.update(0)
    scene.counts()
    scene.enable_timings(True)
    calls = {"set_local_ms": [], "despawn_ms": [], "spawn_ms": [], "update_counts_ms": [], "frame_ms": []}
    k_series, u_series, slow_series = [], [], []
    rec = 0
    for f in range(warm + frames):
        c = f % cohorts
        dead = pin(tmpl[c]["ix"] | (gen[c] << np.uint32(24)))
        gen[c] = (gen[c] + 1) & 0xFF
        fresh, fpar, f_trs, f_aabb, f_mm, f_flags = spawn_batch(c)
        mp = moved_pat[f % 4]
        keep = cohort_of[mp] != c                               # the cohort that was just replaced keeps its spawn TRS
        moved = pin(mp[keep] | (gen[cohort_of[mp[keep]]] << np.uint32(24)))
        mtrs = pin(moved_trs[f % 4][keep])
        D.barrier()
        # game logic first (its 100 MB of pinned TRS cross PCIe while the host does the pool bookkeeping of the streaming
        # calls), then the streaming layer's unloads and loads, then the frame
        t0 = time.perf_counter()
        scene.set_local(moved, mtrs)
        t1 = time.perf_counter()
        scene.despawn(dead)
        t2 = time.perf_counter()
        scene.spawn(fresh, f_trs, fpar, f_aabb, f_mm, f_flags)
        t3 = time.perf_counter()
        scene.update(0)
        if D.world > 1:
            scene.gather_visible(0)
        cnt = scene.counts()
        if D.world > 1:
            scene.synchronize()
        t4 = time.perf_counter()
        if f >= warm:
            k, u = scene.last_timings()
            k_series.append(k)
            u_series.append(u)
            slow_series.append(int(cnt.slowWindows))
            for key, v in zip(calls, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
                calls[key].append(v * 1e3)
            rec = int(cnt.recomputed)
    live, extent = int(cnt.transforms), int(cnt.extent)
    frame_ms = D.max(float(np.median(calls["frame_ms"])))
    k = np.array(k_series)
    first, last = float(k[:10].mean()), float(k[-10:].mean())
    peak, _ = measured_peak()
    alg = ALG_BYTES_DIRTY * rec + ALG_BYTES_CLEAN * (live - rec)
    out = {
        "workload": "BASELINE.json configs[4], per-GPU share: %d instances in depth-4 groups, %d views; per frame ~10 %% despawn "
                    "(whole world sectors, a random tenth of them) + as many spawn (the same sectors in a fresh random order) "
                    "+ 30 %% setLocal (random instances), through the C ABI from pinned host buffers" % (n, args.views),
        "frames": frames, "n_gpus": D.world, "instances_per_gpu": live, "slots_walked": extent,
        "e2e_frame_ms_median": frame_ms, "e2e_value": live * D.world / (frame_ms * 1e-3), "unit": UNIT,
        "calls_ms_median": {k_: float(np.median(v)) for k_, v in calls.items()},
        "device_update_ms_median": float(np.median(u_series)),
        "fused_kernel_ms": {"first10_mean": first, "last10_mean": last, "drift": last / first - 1.0,
                            "min": float(k.min()), "max": float(k.max()), "series": [round(float(x), 4) for x in k]},
        "slow_windows": {"first": slow_series[0], "last": slow_series[-1], "max": max(slow_series)},
        "recomputed_last_frame": rec,
        "roofline": {"bound": "hbm", "achieved": alg / (last * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (last * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(alg),
                     "formula": "132 B x recomputed + 96 B x clean (SURVEY.md 8d), fused kernel, mean of the last 10 frames"},
        "per_frame_bytes_h2d": int(len(dead) * 4 + len(fresh) * (4 + 4 + 36 + 24 + 8 + 4) + len(moved) * 40),
    }
    scene.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=DEFAULT_INSTANCES, help="instances per GPU")
    ap.add_argument("--views", type=int, default=5)
    ap.add_argument("--ref-sample", type=int, default=REF_SAMPLE)
    ap.add_argument("--quick", action="store_true", help="reference arm: a bounded 2 M sample instead of the whole set")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-churn", action="store_true")
    ap.add_argument("--no-partial", action="store_true")
    ap.add_argument("--churn-instances", type=int, default=CHURN_INSTANCES)
    ap.add_argument("--churn-frames", type=int, default=100)
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--sample-instances", type=int, default=REF_SAMPLE, help=argparse.SUPPRESS)
    args = ap.parse_args()

    if args.cpu_worker:
        return cpu_worker(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    # Libraries chat on stdout (NCCL prints "NCCL version ..." there at init): everything written to fd 1 until the
    # result line goes to stderr instead, so that stdout carries exactly ONE JSON line.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import scgpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the scene-update path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    D = Dist(world, local_rank)

    n = args.instances
    sc = build_scene(n, rank, world=world)
    vps = views_for(sc, args.views)
    scene = scgpu.Scene(n, max_views=args.views, device=local_rank, max_entity_index=n)
    entity = np.arange(n, dtype=np.uint32)
    from scgpu import scenes
    scene.spawn(entity, sc["trs9"], scenes.parent_handles(sc["parent"], entity), sc["aabb6"], sc["mesh_mat"], sc["flags"])
    scene.set_views(vps)
    comm_setup(scene, D, rank)

    stream = torch.cuda.ExternalStream(scene.stream, device=torch.device("cuda", local_rank))

    def step():
        scene.mark_all_dirty()
        scene.update(0)
        if world > 1:
            scene.gather_visible(0)

    sampler = ClockSampler(local_rank)  # nvmlInit takes tens of ms and differs per process: keep it out of the region
    sampler.start()
    scene.enable_timings(True)
    for _ in range(args.warmup):
        step()
    D.barrier()

    # ---- timed region: K steps, device-resident inputs, CUDA events on the context stream -----------------
    # the per-kernel events sit between the frame's launches and cut their programmatic-dependent-launch chain, so only
    # every TIMING_EVERY-th step carries them (kernel_ms_avg is the mean over those steps of the timed region)
    scene.enable_timings(True, every=TIMING_EVERY)  # events exist already: this only resets the ring
    sampler.armed = True
    launches0 = scene.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    D.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = scene.launches - launches0
    dev_ms = ev0.elapsed_time(ev1)
    k_ms, u_ms = scene.read_timings(min(args.steps, 256))
    scene.enable_timings(False)
    counts = scene.counts()
    vis_counts = [int(counts.visible[v]) for v in range(args.views)]

    per_rank = None
    if world > 1:
        # per-rank view of the same timed region (diagnostic: which rank sets the max, and is it its kernels or the gather)
        mine = {"rank": rank, "ms_per_step": dev_ms / args.steps, "kernel_ms": float(np.mean(k_ms)) if len(k_ms) else None,
                "update_ms": float(np.mean(u_ms)) if len(u_ms) else None, "sm_mhz": clocks.get("sm_mhz")}
        per_rank = D.gather_objects(mine)
    ms_per_step = D.max(dev_ms) / args.steps
    total_instances = n * world
    value = total_instances / (ms_per_step * 1e-3)

    gather = check_gather(scene, D, rank, args.views)

    # ---- e2e: the same frame through the C ABI with HOST buffers: upload of every instance's new local transform,
    #      update, read-back of counts and of every view's visible list (N > 1: the GATHERED lists, on the root) -----
    e2e = None
    if not args.no_e2e:
        pin_e = torch.from_numpy(entity).pin_memory()
        pin_t = torch.from_numpy(sc["trs9"]).pin_memory()
        pin_pr = torch.from_numpy(np.ascontiguousarray(sc["trs9"][:, 0:6])).pin_memory()
        pe, pt, ppr = pin_e.numpy(), pin_t.numpy(), pin_pr.numpy()
        cap = min(n * (world if rank == 0 else 1), 8 << 20)   # per view; the lists are ~1 % of the scene
        out_lists = [torch.empty(cap, dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(args.views)]

        def read_back():
            scene.counts()
            got = 0
            if world > 1:
                scene.gather_visible(0)
                if rank == 0:
                    for v in range(args.views):
                        got += len(scene.read_gathered_visible(v, out_lists[v]))
                else:
                    scene.synchronize()   # this rank's lists have left for the root
            else:
                for v in range(args.views):
                    got += len(scene.read_visible(v, out_lists[v]))
            return got

        uploads = {
            # what the engine's per-frame writers change: position + rotation (physics sync sc_physics.cpp:1178-1184, traffic
            # sc_traffic_ai.cpp:449-457), whole pool in dense order, no handles: 24 B per instance
            "pos_rot_range": (lambda: scene.set_local_range(0, ppr, 6), n * 24),
            # every field of every Transform, dense order: 36 B per instance
            "trs_range": (lambda: scene.set_local_range(0, pt, 9), n * 36),
            # round 1's form: setLocal by entity handle, 4 + 36 B per instance
            "trs_by_handle": (lambda: scene.set_local(pe, pt), n * 40),
        }
        e2e_steps = max(3, min(args.steps, 10))
        variants = {}
        for name, (upload, h2d) in uploads.items():
            def e2e_step():
                upload()
                scene.update(0)
                return read_back()
            e2e_step()
            D.barrier()
            t0 = time.perf_counter()
            got = 0
            for _ in range(e2e_steps):
                got = e2e_step()
            D.barrier()
            sec = D.max((time.perf_counter() - t0) / e2e_steps)
            variants[name] = {"value": total_instances / sec, "ms_per_step": sec * 1e3, "h2d_bytes_per_step": int(h2d),
                              "d2h_bytes_per_step": int(got * 4 + 4 * (scgpu.MAX_VIEWS + 3))}
        head = variants["pos_rot_range"]
        e2e = {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": head["h2d_bytes_per_step"],
               "d2h_bytes_per_step": head["d2h_bytes_per_step"], "ms_per_step": head["ms_per_step"], "steps": e2e_steps,
               "path": "scgpuSetLocalRange(whole pool, pinned host position+rotation: the fields the engine's per-frame writers "
                       "change) + scgpuUpdate + scgpuGetCounts + " +
                       ("scgpuGatherVisible + scgpuReadGatheredVisible x views on rank 0" if world > 1 else "scgpuReadVisible x views"),
               "variants": variants}

    partial = clean = None
    if not args.no_partial:
        clean = clean_leg(scene, n, args, D, stream, torch)
        partial = partial_dirty_leg(scene, sc, n, args, D, stream, torch)
    scene.close()
    del scene

    churn = None
    if not args.no_churn:
        churn = churn_leg(args, D, rank, local_rank, torch)

    if rank == 0:
        peak, peak_src = measured_peak()
        k_avg_ms = float(np.mean(k_ms)) if len(k_ms) else None
        alg_bytes = ALG_BYTES_DIRTY * n
        achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms else None
        traffic = recorded_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm", "kernel": "k_update_win<V> (fused transform + sphere + V-view cull + visibility bits; hierarchy windows per warp)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_instance": ALG_BYTES_DIRTY, "kernel_ms_avg": k_avg_ms,
                "kernel_share_of_step": (k_avg_ms / ms_per_step) if k_avg_ms else None,
                "kernel_samples": int(len(k_ms)), "kernel_events_every_n_steps": TIMING_EVERY,
                "update_ms_avg": float(np.mean(u_ms)) if len(u_ms) else None,
                "step_frac_of_peak": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                "traffic_source": (traffic or {}).get("source"),
            },
            "visible_per_view": vis_counts, "wall_ms_per_step": t_wall * 1e3 / args.steps,
        }
        if gather:
            line["gather"] = gather
        if clean:
            line["clean_frame"] = clean
        if partial:
            line["partial_dirty"] = partial
        if churn:
            line["churn"] = churn
        if per_rank:
            line["per_rank"] = per_rank
        if world == 1 and not args.no_cpu_baseline:
            res = run_cpu_worker(args.ref_sample, args.views, 3, 1)
            if "error" in res:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "failed: " + res["error"][:160]}
            else:
                line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": res.get("threads", res["cores"]),
                                        "kind": res["kind"], "sample": res["sample"], "host_cpus": res["cores"],
                                        "stages_ms": {k: res[k] for k in ("transform_ms", "cull_ms", "prep_ms") if k in res}}
        sys.stdout.flush()
        os.dup2(result_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    D.close()


def _guarded_main():
    """one rank that dies must not leave the others waiting in a collective until the driver's limit: any exception ends
    the process at once, and a watchdog ends a run that exceeds SCGPU_BENCH_LIMIT_S (default 20 min)"""
    limit = float(os.environ.get("SCGPU_BENCH_LIMIT_S", "1200"))
    watchdog = threading.Timer(limit, lambda: (sys.stderr.write("bench.py: time limit exceeded\n"), os._exit(3)))
    watchdog.daemon = True
    watchdog.start()
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    _guarded_main()
