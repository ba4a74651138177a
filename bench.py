#!/usr/bin/env python
"""bench.py — scene-update hot path benchmark (transform -> cull (V views) -> compact), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (libscgpu.so through the C ABI)
  python bench.py --impl reference ...                          the reference's own CPU code (oracle/_ref), rank 0 only

A "step" is one frame over the whole per-GPU instance set with EVERY instance dirty (worst case: each one is
re-transformed, culled against all views and compacted). Workload at any N: the north-star per-GPU share —
16 Mi instances per GPU in depth-4 groups, main camera + 4 shadow cascades (BASELINE.json configs[2]; 8 GPUs of
it = configs[3], 128 Mi instances sharded by world cell). Weak scaling: every rank owns a contiguous block of
world cells; the only exchange is the gather of the compacted lists to rank 0 (inside the timed step for N>1).
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
REF_SAMPLE = 2_000_000  # instances per step of the CPU reference arm (bounded sample)


def build_scene(n, rank, seed=424242):
    from scgpu import scenes
    if os.environ.get("SCGPU_BENCH_WORKLOAD", "hier") == "flat":  # diagnostic only; the reported workload is "hier"
        sc = scenes.city_flat(n, seed=seed + 7919 * rank)
        return sc
    sc = scenes.city_hier(n, seed=seed + 7919 * rank)
    # rank r owns the block of world cells shifted by r grid sides along +x: contiguous cell blocks of one world
    shift = np.float32(rank * sc["side"] * scenes.SECTOR_SIZE)
    roots = sc["parent"] < 0
    sc["trs9"][roots, 0] += shift
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

def cpu_worker(args):
    """Times the reference's own TransformSystem -> CullingSystem x V -> RenderPrepStreamingSystem on a bounded
    sample of the bench workload. Prints one JSON object per step set."""
    sys.path.insert(0, str(ROOT / "tests"))
    import ctypes as C
    import oracle_bind
    n = args.sample_instances
    sc = build_scene(n, 0)
    vps = views_for(sc, args.views)
    from scgpu import scenes
    out = {"cores": os.cpu_count(), "sample": f"{n} instances of the same depth-4 city, {args.views} views, all dirty, per step"}
    if oracle_bind.ref_available():
        r = oracle_bind.RefScene(0)
        e = r.create_entities(n)
        par = scenes.parent_handles(sc["parent"], e)
        r.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        L = r.L
        vp = np.ascontiguousarray(vps, np.float32)
        ee = np.ascontiguousarray(e)
        tT, tC, tP = C.c_double(), C.c_double(), C.c_double()
        per_step = []
        total = args.warmup + args.steps
        for it in range(total):
            L.screfTimeFrame(r.w, 1, args.views, vp.ctypes.data_as(C.c_void_p), n, ee.ctypes.data_as(C.c_void_p), 0,
                             C.byref(tT), C.byref(tC), C.byref(tP))
            if it >= args.warmup:
                per_step.append((tT.value, tC.value, tP.value))
        a = np.array(per_step)
        out.update(kind="reference", threads=int(L.screfJobWorkers()) + 1,
                   transform_ms=float(a[:, 0].mean() * 1e3), cull_ms=float(a[:, 1].mean() * 1e3),
                   prep_ms=float(a[:, 2].mean() * 1e3))
        sec = float(a.sum(axis=1).mean())
        r.close()
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
    if rank != 0:
        return
    res = run_cpu_worker(args.ref_sample, args.views, args.steps, args.warmup, timeout=3000)
    if "error" in res:
        print(json.dumps({"impl": "reference", "unavailable": res["error"].replace("\n", " ")[:200]}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, res["sample_instances"], sample=True),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res.get("threads", res["cores"]),
                         "kind": res["kind"], "sample": res["sample"],
                         "stages_ms": {k: res[k] for k in ("transform_ms", "cull_ms", "prep_ms") if k in res},
                         "host_cpus": res["cores"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_per_gpu, sample=False):
    return {
        "workload": "synthetic city, depth-4 groups (vehicles+wheels, peds+attachments), main view + 4 shadow "
                    "cascades, all instances dirty every step (BASELINE.json configs[2]; x8 GPUs = configs[3])",
        "instances_per_gpu": int(n_per_gpu), "views": args.views, "dirty_fraction": 1.0,
        "sharding": ("world cell blocks, one process per GPU; visible lists gathered to rank 0 through "
                     + ("NVLink peer memory" if os.environ.get("SCGPU_GATHER", "peer") == "peer" else "NCCL send/recv"))
        if args.gpus > 1 else "single GPU",
        "l2": "inputs (>= 2 GB per step) exceed the 126 MB L2; no explicit flush" if not sample else "cpu sample",
    }


# --------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=DEFAULT_INSTANCES, help="instances per GPU")
    ap.add_argument("--views", type=int, default=5)
    ap.add_argument("--ref-sample", type=int, default=REF_SAMPLE)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    import torch.distributed as dist
    import scgpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the scene-update path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.instances
    sc = build_scene(n, rank)
    vps = views_for(sc, args.views)
    scene = scgpu.Scene(n, max_views=args.views, device=local_rank, max_entity_index=n)
    entity = np.arange(n, dtype=np.uint32)
    from scgpu import scenes
    scene.spawn(entity, sc["trs9"], scenes.parent_handles(sc["parent"], entity), sc["aabb6"], sc["mesh_mat"], sc["flags"])
    scene.set_views(vps)
    if world > 1:
        uid = [scgpu.Scene.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        scene.comm_init(world, rank, uid[0])
        if os.environ.get("SCGPU_GATHER", "peer") == "peer":  # lists travel through NVLink peer memory, not NCCL
            scene.enable_peer_gather(0)

    stream = torch.cuda.ExternalStream(scene.stream, device=torch.device("cuda", local_rank))

    def step():
        scene.mark_all_dirty()
        scene.update(0)
        if world > 1:
            scene.gather_visible(0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)  # nvmlInit takes tens of ms and differs per process: keep it out of the region
    sampler.start()
    scene.enable_timings(True)
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: K steps, device-resident inputs, CUDA events on the context stream -----------------
    scene.enable_timings(True)  # events exist already: this only resets the ring
    sampler.armed = True
    launches0 = scene.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = scene.launches - launches0
    dev_ms = ev0.elapsed_time(ev1)
    k_ms, u_ms = scene.read_timings(min(args.steps, 256))
    scene.enable_timings(False)
    counts = scene.counts()
    vis_counts = [int(counts.visible[v]) for v in range(args.views)]

    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # per-rank view of the same timed region (diagnostic: which rank sets the max, and is it its kernels or the gather)
        mine = {"rank": rank, "ms_per_step": dev_ms / args.steps, "kernel_ms": float(np.mean(k_ms)) if len(k_ms) else None,
                "update_ms": float(np.mean(u_ms)) if len(u_ms) else None, "sm_mhz": clocks.get("sm_mhz")}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    dev_ms_max = float(t.item())
    ms_per_step = dev_ms_max / args.steps
    total_instances = n * world
    value = total_instances / (ms_per_step * 1e-3)

    # ---- e2e: the same frame through the C ABI with HOST buffers: upload of every instance's TRS, update,
    #      read-back of counts and of every view's visible list ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        pin_e = torch.from_numpy(entity).pin_memory()
        pin_t = torch.from_numpy(sc["trs9"]).pin_memory()
        pe, pt = pin_e.numpy(), pin_t.numpy()
        out_lists = [torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(args.views)]

        def e2e_step():
            scene.set_local(pe, pt)
            scene.update(0)
            if world > 1:
                scene.gather_visible(0)
            scene.counts()
            got = 0
            for v in range(args.views):
                got += len(scene.read_visible(v, out_lists[v]))
            return got

        e2e_steps = max(3, min(args.steps, 10))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        got = 0
        for _ in range(e2e_steps):
            got = e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
        e2e = {"value": total_instances / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(n * (4 + 36)),
               "d2h_bytes_per_step": int(got * 4 + 4 * (scgpu.MAX_VIEWS + 2)), "ms_per_step": e2e_s * 1e3,
               "steps": e2e_steps,
               "path": "scgpuSetLocal(all, pinned host TRS) + scgpuUpdate + scgpuGetCounts + scgpuReadVisible x views"}

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
                "bound": "hbm", "kernel": "k_update_win<V> (fused transform + sphere + V-view cull + tile counts; hierarchy windows per warp)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_instance": ALG_BYTES_DIRTY, "kernel_ms_avg": k_avg_ms,
                "kernel_share_of_step": (k_avg_ms / ms_per_step) if k_avg_ms else None,
                "update_ms_avg": float(np.mean(u_ms)) if len(u_ms) else None,
                "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                "traffic_source": (traffic or {}).get("source"),
            },
            "visible_per_view": vis_counts, "wall_ms_per_step": t_wall * 1e3 / args.steps,
        }
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

    scene.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
