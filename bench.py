#!/usr/bin/env python
"""bench.py -- depth-frames/sec into a depth-16 SVO (+ raycast Mrays/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one 640x480 RGB-D frame of the synthetic orbit (scene S0) integrated into ONE
incremental depth-16 SVO through the C ABI.  `value` = frames/s with the frames already resident in HBM (a ring of
distinct frames larger than L2); `e2e` = the same through osl_integrate_depth_host with pinned HOST frames (H2D inside
the timed region, the per-frame result block read back).  N > 1: one process per GPU, every rank fuses its OWN stream
into its OWN map (replicas, no data-path collective -> "weak" scaling); time = max over ranks.

--impl reference times the reference's own path (its host code + its original CUDA kernels re-targeted to sm_100a,
oracle/_ref, "ref + 64-bit patch" because depth 16 > 10) on the same config; if that library is missing it falls back
to the single-threaded CPU oracle port."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Deployment knob of the frame pipeline: with CUDA's default of 8
# hardware work queues every one of the pipeline's streams owns a queue and the GPU front end spends ~5 us per frame
# switching between them; 3 queues measured best, alone (32.6 k -> 35.7 k frames/s) and under torchrun with NCCL
# initialised (2 GPUs: 67.9 k -> 70.4 k).  Must be set before CUDA initialises; an explicit setting wins.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "3")

W, H, DEPTH = 640, 480, 16
RING = 136           # distinct frames in HBM: 136 * 1.536 MB = 209 MB > 126 MB L2
RAY_W, RAY_H = 640, 480
FOV = 45.0
METRIC = "depth_frames_per_sec_640x480_into_depth16_svo"
# one `ncu --set full` capture of k_raycast on the bench map (profiles/r02_ncu_full_summary_raycast.csv)
RAY_NCU = {"dram_bytes": 414976,
           "issue": {"issue_slots_active_pct": 72.6, "inst_per_cycle_per_sm": 2.38, "peak_inst_per_cycle_per_sm": 4.0,
                     "active_threads_per_warp_inst": 22.45, "l1_hit_pct": 97.1, "warps_active_pct": 32.9,
                     "top_stalls": "wait 29 %, not selected 17 %, selected 15 %, long scoreboard 11 % "
                                   "(profiles/r02_ncu_raycast_stalls.csv)"},
           "source": "profiles/r02_ncu_full_summary_raycast.csv (640x480, 25-frame bench map, 198 us under ncu)"}


def make_ring(synth, n_frames, seed0=0):
    depths, rgbs, poses = [], [], []
    for k in range(n_frames):
        pose = synth.orbit_pose(k)
        d, c = synth.make_frame(W, H, pose, seed=seed0 + k)
        depths.append(d)
        rgbs.append(c)
        poses.append(pose)
    return np.stack(depths), np.stack(rgbs), poses


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def bind_to_gpu_numa_node(local):
    """Pin this rank to the CPU cores of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated
    (first-touch puts the frame ring next to the GPU's PCIe root): with 8 ranks the H2D copies of the e2e path
    otherwise all pull from one socket.  Best effort: silently does nothing when the topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind_to_gpu_numa_node(local)
    import __graft_entry__ as graft
    graft.wait_for_cuda(device=local)  # (a fresh box sometimes refuses the first CUDA initialisation for a few seconds)
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_ours(args):
    import torch
    import __graft_entry__ as graft
    pkg = graft.load_package()
    lib = pkg.lib()
    rank, world, local = dist_setup(args.gpus)
    K, Wm = args.steps, args.warmup
    center, half = pkg.synth.tree_params(DEPTH)
    fx, fy = pkg.synth.focal(W, H)

    # every rank fuses its own stream (different noise seeds) into its own map
    depths, rgbs, poses = make_ring(pkg.synth, RING, seed0=1000 * rank)
    d_depth = torch.from_numpy(depths).cuda()
    d_rgb = torch.from_numpy(rgbs).cuda()
    h_depth = torch.from_numpy(depths).pin_memory()
    h_rgb = torch.from_numpy(rgbs).pin_memory()
    pose_c = [pkg.capi._f(pkg.capi.mat_colmajor(p)) for p in poses]
    stream = torch.cuda.Stream()
    sp = stream.cuda_stream

    # raw pointers of every ring slot, taken once: the timed loops must not pay for tensor indexing
    dptr = [(d_depth[j].data_ptr(), d_rgb[j].data_ptr()) for j in range(RING)]
    hptr = [(h_depth[j].data_ptr(), h_rgb[j].data_ptr()) for j in range(RING)]
    f_dev, f_host = lib.osl_integrate_depth, lib.osl_integrate_depth_host

    def integrate_resident(svo, k):
        j = k % RING
        rc = f_dev(svo._h, dptr[j][0], dptr[j][1], W, H, fx, fy, pose_c[j], sp)
        if rc:
            raise RuntimeError("osl_integrate_depth -> %d" % rc)

    def integrate_host(svo, k):
        j = k % RING
        rc = f_host(svo._h, hptr[j][0], hptr[j][1], W, H, fx, fy, pose_c[j], sp)
        if rc:
            raise RuntimeError("osl_integrate_depth_host -> %d" % rc)

    def timed(fn, svo):
        for k in range(Wm):
            fn(svo, k)
        barrier(world)
        sampler = ClockSampler(local) if rank == 0 else None
        l0 = lib.osl_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0 = svo.counters().total_algorithmic_bytes  # waits for the warm-up frames
        barrier(world)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for k in range(Wm, Wm + K):
                fn(svo, k)  # asynchronous: the host only throttles when it is > 3 frames ahead
            svo.join(sp)  # pipelined frames finish on the library's streams: order the timing event after them
            e1.record(stream)
        barrier(world)
        ms = e0.elapsed_time(e1)
        bytes_alg = svo.counters().total_algorithmic_bytes - b0
        clocks = sampler.stop() if sampler else None
        return max_over_ranks(ms, world), lib.osl_launch_count() - l0, bytes_alg, clocks

    # resident inputs are complete before the timed region starts -> pipelined mode (osl_svo_set_pipeline)
    svo = pkg.SVO(center, half, DEPTH, reserve_nodes=1 << 24, device=local).set_pipeline(True)
    ms, launches, bytes_alg, clocks = timed(integrate_resident, svo)
    nodes = svo.size
    last = svo.counters()
    # per-kernel CUDA-event times (non-pipelined frames on the same map), for the roofline of the dominant kernel
    svo.set_pipeline(False).set_stage_timing(True)
    stage = np.zeros(4)
    for k in range(Wm + K, Wm + K + 20):
        integrate_resident(svo, k)
        stage += np.array(svo.stage_times())
    stage /= 20.0
    svo.set_stage_timing(False)
    # raycast of the fused map from the last camera pose: image rows dealt to the ranks in interleaved bands; at N > 1
    # rank 0's tree is first replicated to every rank (NCCL broadcast of the flat pool), time = max over ranks
    view = (np.diag([-1.0, 1.0, -1.0, 1.0]) @ np.linalg.inv(poses[(Wm + K - 1) % RING].astype(np.float64))).astype(np.float32)
    if world > 1:
        torch.cuda.synchronize()
        pkg.shard.replicate_tree(svo, 0, device="cuda")
    band = pkg.shard.band_height(RAY_H, world)
    my_rows = sum(r for _, r in pkg.shard.row_bands(RAY_H, world, rank, band))
    out_rows = torch.empty((max(my_rows, 1), RAY_W, 4), dtype=torch.uint8, device="cuda")

    def render():  # one launch: all the interleaved bands this rank owns
        svo.raycast_bands(out_rows, RAY_W, RAY_H, band, world, rank, FOV, view, stream=sp)

    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        render()
    barrier(world)
    with torch.cuda.stream(stream):
        r0.record(stream)
        for _ in range(10):
            render()
        r1.record(stream)
    barrier(world)
    ray_ms = max_over_ranks(r0.elapsed_time(r1) / 10, world)
    st = pkg.RaycastStats()
    svo.raycast(RAY_W, RAY_H, FOV, view, stats=st)
    # the same map at configs[1]'s render size (1920x1080), rank 0's share of the bands
    hd_rows = sum(r for _, r in pkg.shard.row_bands(1080, world, rank, pkg.shard.band_height(1080, world)))
    out_hd = torch.empty((max(hd_rows, 1), 1920, 4), dtype=torch.uint8, device="cuda")
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        svo.raycast_bands(out_hd, 1920, 1080, pkg.shard.band_height(1080, world), world, rank, FOV, view, stream=sp)
    barrier(world)
    with torch.cuda.stream(stream):
        h0.record(stream)
        for _ in range(5):
            svo.raycast_bands(out_hd, 1920, 1080, pkg.shard.band_height(1080, world), world, rank, FOV, view, stream=sp)
        h1.record(stream)
    barrier(world)
    hd_ms = max_over_ranks(h0.elapsed_time(h1) / 5, world)
    svo.close()

    svo2 = pkg.SVO(center, half, DEPTH, reserve_nodes=1 << 24, device=local)
    ms_e2e, _, _, _ = timed(integrate_host, svo2)
    svo2.close()

    # camera tracking (the step before integration): frame-to-frame ICP on resident depth frames,
    # 27 launches per frame, pose read back once at the end
    cam = pkg.RGBDCamera(W, H, (fx, fy), exact_jacobian=True, device=local)
    n_trk = min(K, 100)
    for k in range(3):
        lib.osl_tracker_update(cam._h, dptr[k % RING][0], sp)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.osl_launch_count()
    with torch.cuda.stream(stream):
        c0.record(stream)
        for k in range(3, 3 + n_trk):
            lib.osl_tracker_update(cam._h, dptr[k % RING][0], sp)
        c1.record(stream)
    torch.cuda.synchronize()
    trk_ms = max_over_ranks(c0.elapsed_time(c1) / n_trk, world)
    trk_launches = (lib.osl_launch_count() - l0) / n_trk
    trk_lost = cam.lost
    # the whole SLAM frame of main.cpp:33-44 with tracking live: tracker, then integration with the pose read on the
    # device (osl_integrate_depth_posed) -- 31 launches, no host round trip
    svo3 = pkg.SVO(center, half, DEPTH, reserve_nodes=1 << 24, device=local).set_pipeline(True)
    d_pose = cam.pose_device()

    def slam_frame(k):
        j = k % RING
        lib.osl_tracker_update(cam._h, dptr[j][0], sp)
        rc = lib.osl_integrate_depth_posed(svo3._h, dptr[j][0], dptr[j][1], W, H, fx, fy, d_pose, sp)
        if rc:
            raise RuntimeError("osl_integrate_depth_posed -> %d" % rc)

    for k in range(5):
        slam_frame(k)
    svo3.counters()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        s0.record(stream)
        for k in range(5, 5 + n_trk):
            slam_frame(k)
        svo3.join(sp)
        s1.record(stream)
    torch.cuda.synchronize()
    slam_ms = max_over_ranks(s0.elapsed_time(s1) / n_trk, world)
    svo3.close()
    del cam

    ray_alg_gbs = (4 * st.rays + 4 * st.visits + 4 * st.steps) / (ray_ms / 1e3) / 1e9
    total_frames = sum_over_ranks(float(K), world)
    value = total_frames / (ms / 1e3)
    e2e_value = total_frames / (ms_e2e / 1e3)
    peak, peak_src = measured_peak()
    achieved = bytes_alg / (ms / 1e3) / 1e9  # GB/s of algorithmic bytes over the integrate pipeline of this rank
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 keys / u32 nodes / f32 geometry", "data": "synthetic",
        "config": {"workload": "cfg3/4-style: 640x480 synthetic RGB-D orbit, incremental fusion into one depth-16 SVO "
                               "(1 cm leaves, half edge 655.36 m) per GPU",
                   "l2": "inputs cycle through a %d-frame ring (%.0f MB > 126 MB L2)" % (RING, RING * W * H * 5 / 1e6),
                   "multi_gpu": "replicas only: one independent stream+map per rank, no data-path collective",
                   "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"),
                   "nodes_after": nodes},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": W * H * 5,
                "d2h_bytes_per_step": int(lib.osl_frame_result_bytes()),
                "api": "osl_integrate_depth_host: pinned host frames, H2D on the library's copy stream, per-frame "
                       "result block written by the device into pinned host memory"},
        "gpu_launches": int(launches),
        "raycast": {"mrays_per_s": RAY_W * RAY_H / (ray_ms / 1e3) / 1e6, "ms": ray_ms, "res": [RAY_W, RAY_H],
                    "mode": "ref_exact", "rows": "interleaved bands over %d rank(s)" % world,
                    "steps_per_ray": st.steps / float(st.rays),
                    "algorithmic_gbs": (4 * st.rays + 4 * st.visits + 4 * st.steps) / (ray_ms / 1e3) / 1e9,
                    "at_1920x1080": {"ms": hd_ms, "mrays_per_s": 1920 * 1080 / (hd_ms / 1e3) / 1e6},
                    # B_ray = 4R + sum over steps of (4 v + 4): what the reference's root-to-LOD descents read;
                    # our descents resume at cached ancestors and hit L1 (97 %), so the DRAM side of it is tiny
                    "roofline": {"bound": "hbm", "achieved": ray_alg_gbs, "peak": peak, "unit": "GB/s",
                                 "frac": (ray_alg_gbs / peak) if ray_alg_gbs else None,
                                 "traffic": RAY_NCU["dram_bytes"], "traffic_source": RAY_NCU["source"],
                                 "issue": RAY_NCU["issue"],
                                 "note": "algorithmic bytes are L1 hits: the kernel is bound by instruction issue and "
                                         "dependent-load latency, not by HBM (DESIGN.md section 6)"}},
        "tracking": {"frames_per_s": 1e3 / trk_ms * world, "ms_per_frame": trk_ms, "launches_per_frame": trk_launches,
                     "lost": bool(trk_lost), "slam_frames_per_s": 1e3 / slam_ms * world, "slam_ms_per_frame": slam_ms,
                     "what": "sensor::RGBDCamera::update: bilateral + 3-level pyramid + "
                     "19 ICP iterations, device-side solve, depth frames resident"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of the four kernels of one frame, one ncu --set full
                     # capture (profiles/r01_ncu_full_summary_v07.csv; cold caches: ncu flushes L2 between kernels)
                     "traffic": 2055168, "traffic_source": "profiles/r02_ncu_full_summary_kframe_v2.csv (one k_frame launch)",
                     "peak_source": peak_src,
                     "kernel": "k_frame: ONE launch per frame whose four roles (emit / sort / structure / values) each "
                               "work on the frame that has reached them; achieved = B_int of one frame / average "
                               "launch period over the timed region (CUDA events around the PDL-chained launches), "
                               "B_int = 5N+8U+68S+68*sum(P_l)",
                     "launches_in_timed_region": int(launches),
                     "bytes_per_frame": bytes_alg / float(K),
                     "counters_last_frame": {"N": int(last.n_points), "V": int(last.n_valid), "U": int(last.n_unique),
                                             "S": int(last.n_split),
                                             "sum_P": int(sum(list(last.parents)[:DEPTH]))},
                     "stage_us_unpipelined": {"k_emit": 1e3 * stage[0], "k_sort": 1e3 * stage[1],
                                              "k_structure": 1e3 * stage[2], "k_levels": 1e3 * stage[3]},
                     "note": "latency-bound: a frame's algorithmic bytes (~1.9 MB) take 0.3 us at HBM speed; see "
                             "DESIGN.md section 3"},
        "clocks": clocks,
    }
    # BASELINE configs 2 and 5 on their named inputs (tools/cfg_mesh_bench.py; the meshes are staged into
    # baseline/_assets by __graft_entry__.build()): voxelise -> svoFromVoxelGrid -> raycast, every rank its own replica
    # of the map, the image in interleaved row bands over the ranks
    if not args.no_mesh_configs:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import cfg_mesh_bench
        for name in ("cfg2", "cfg5"):
            try:
                res, rms = cfg_mesh_bench.run(name, None, 3, rank, world, local)
                rms = max_over_ranks(rms, world)
                Wm_, Hm_ = res["raycast"]["res"]
                res["raycast"]["ms_max_over_ranks"] = rms
                res["raycast"]["mrays_per_s"] = Wm_ * Hm_ / (rms / 1e3) / 1e6
                line[name] = res
            except Exception as exc:  # (a missing asset or an out-of-memory box must not cost the headline line)
                line[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
                max_over_ranks(0.0, world)
    if rank == 0:
        line["cpu_baseline"] = cpu_baseline_port(pkg, depths, rgbs, poses, fx, fy, center, half)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline_port(pkg, depths, rgbs, poses, fx, fy, center, half, budget_s=12.0):
    """The CPU oracle port (1 thread) on a bounded sample of the same workload."""
    from oracle import oracle as orc
    t = orc.OracleSVO(center, half, DEPTH)
    n, t0 = 0, time.time()
    while n < 40 and (time.time() - t0) < budget_s:
        t.integrate_depth(depths[n % RING], rgbs[n % RING], fx, fy, poses[n % RING])
        n += 1
    dt = time.time() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": "first %d frames of the same orbit, oracle/osl_oracle.c single thread" % n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as graft
    pkg = graft.load_package()
    from oracle import ref as R
    K, Wm = args.steps, args.warmup
    center, half = pkg.synth.tree_params(DEPTH)
    fx, fy = pkg.synth.focal(W, H)
    ring = min(RING, max(8, K + Wm))
    depths, rgbs, poses = make_ring(pkg.synth, ring)
    base = {"metric": METRIC, "unit": "frames/s", "n_gpus": args.gpus, "steps": K, "warmup": Wm, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / u32 nodes / f32 geometry",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "cfg3/4-style: 640x480 synthetic RGB-D orbit, incremental fusion into one depth-16 "
                                   "SVO (1 cm leaves, half edge 655.36 m)",
                       "multi_gpu": "the reference has no multi-GPU path: rank 0 runs it on one GPU whatever --gpus says"}}
    have_gpu_ref = False
    try:
        import torch
        have_gpu_ref = R.available(True) and torch.cuda.is_available()
    except Exception:
        have_gpu_ref = False
    if have_gpu_ref:
        import torch
        torch.cuda.set_device(0)
        d_depth = torch.from_numpy(depths).cuda()
        d_rgb = torch.from_numpy(rgbs).cuda()
        t = R.RefSVO(center, half, DEPTH, patched64=True)
        for k in range(Wm):
            t.integrate_depth_dev(d_depth[k % ring].data_ptr(), d_rgb[k % ring].data_ptr(), W, H, fx, fy, poses[k % ring])
        torch.cuda.synchronize()
        t0 = time.time()
        per_frame = []  # the shim's own wall clock around each frame (sync on both sides)
        for k in range(Wm, Wm + K):
            per_frame.append(t.integrate_depth_dev(d_depth[k % ring].data_ptr(), d_rgb[k % ring].data_ptr(), W, H, fx, fy,
                                                   poses[k % ring]))
        torch.cuda.synchronize()
        dt = time.time() - t0
        # end to end from host frames (the reference's readFrame H2D + the same path)
        t2 = R.RefSVO(center, half, DEPTH, patched64=True)
        for k in range(Wm):
            t2.integrate_depth(depths[k % ring], rgbs[k % ring], fx, fy, poses[k % ring])
        t0 = time.time()
        for k in range(Wm, Wm + K):
            t2.integrate_depth(depths[k % ring], rgbs[k % ring], fx, fy, poses[k % ring])
        torch.cuda.synchronize()
        dt2 = time.time() - t0
        view = (np.diag([-1.0, 1.0, -1.0, 1.0]) @ np.linalg.inv(poses[(Wm + K - 1) % ring].astype(np.float64))).astype(np.float32)
        _, ray_ms = t.raycast(RAY_W, RAY_H, FOV, view, want_image=False)
        _, ray_ms = t.raycast(RAY_W, RAY_H, FOV, view, want_image=False)
        trk = R.RefTracker(W, H, fx, fy)
        n_trk = min(K, 20)
        for k in range(3):
            trk.update(depths[k % ring])
        trk_ms = sum(trk.update(depths[k % ring]) for k in range(3, 3 + n_trk)) / n_trk
        base["tracking"] = {"frames_per_s": 1e3 / trk_ms, "ms_per_frame": trk_ms,
                            "what": "the reference's own RGBDCamera::update (rgbd_camera.cpp + its CUDA kernels)"}
        value = K / dt
        base.update({"value": value, "ms_per_step": dt / K * 1e3,
                     # the reference's path is bound by its own cudaMalloc / cudaFree / sync calls, whose cost varies
                     # between boxes and between frames: the per-frame median and minimum give the ratio a stable reading
                     "ms_per_step_median": float(np.median(per_frame)), "ms_per_step_min": float(np.min(per_frame)),
                     "value_from_median": 1e3 / float(np.median(per_frame)),
                     "value_from_min": 1e3 / float(np.min(per_frame)),
                     "cpu_baseline": {"value": value, "unit": "frames/s", "cores": 1, "kind": "reference",
                                      "sample": "%d frames; reference host code on 1 host thread + its ORIGINAL CUDA "
                                                "kernels (svo.cu, image_kernels.cu) re-targeted to sm_100a on cuda:0, "
                                                "'ref + 64-bit patch' because depth 16 > 10; nproc=%d"
                                                % (K, os.cpu_count())},
                     "e2e": {"value": K / dt2, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "raycast": {"mrays_per_s": RAY_W * RAY_H / (ray_ms / 1e3) / 1e6, "ms": ray_ms}})
    else:
        from oracle import oracle as orc
        t = orc.OracleSVO(center, half, DEPTH)
        for k in range(min(Wm, 2)):
            t.integrate_depth(depths[k % ring], rgbs[k % ring], fx, fy, poses[k % ring])
        n, t0 = 0, time.time()
        while n < K and time.time() - t0 < 60.0:
            t.integrate_depth(depths[(Wm + n) % ring], rgbs[(Wm + n) % ring], fx, fy, poses[(Wm + n) % ring])
            n += 1
        dt = time.time() - t0
        value = n / dt
        base.update({"value": value, "ms_per_step": dt / n * 1e3,
                     "cpu_baseline": {"value": value, "unit": "frames/s", "cores": 1, "kind": "port",
                                      "sample": "%d frames, oracle/osl_oracle.c single thread (the reference has no CPU "
                                                "path and oracle/_ref is not available here)" % n},
                     "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-mesh-configs", action="store_true", help="skip the cfg2 / cfg5 mesh objects of the line")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
