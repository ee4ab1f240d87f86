#!/usr/bin/env python
"""bench.py — Mpoints/s of the V-PCC patch-generation + image-formation hot path (SURVEY.md §8a rows a1–a26).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F] [--scale S]

One "step" = one GOF (--frames F frames per GPU) of synthetic longdress-like frames (tests/synth.py `figure`, 10-bit, ≈0.83 M points
per frame at the default scale) pushed through the whole hot path: generateSegments, placeSegments, occupancy / geometry
image formation, generatePointCloud, colour transfer, attribute image formation and padding (PCCEncoder.cpp:103-424,
the three videoEncoder.compress calls excluded, occupancy/geometry video treated as lossless).  No dataset ships with
the reference, so the data is synthetic and says so.

Steps are independent GOFs; --gofs-in-flight of them are processed concurrently (one library context each, started staggered):
the orientation walk of a frame is a 1-2 s single-warp latency chain, so the GPU is kept busy by the data-parallel stages of
the other GOFs in flight. value / e2e are whole-job throughputs over the timed region (K GOFs, barrier to barrier).

Multi-GPU (torchrun, one process per GPU): the frames of every GOF are sharded over the ranks (rank r takes F frames of
an N*F-frame GOF: weak scaling); the one cross-frame coupling of the all-intra path — the common canvas size — is one
NCCL all-reduce(MAX) of two integers between packing and image formation.  Timing: barrier + synchronize on both sides,
max over ranks.

JSON keys follow the driver contract.  value = throughput over the DEVICE window (first compute span to last span of any
frame stream, CUDA events; input already uploaded), e2e = wall clock through the C ABI with host buffers including the
H2D of the clouds and the D2H of every frame handed to the video codec, roofline = dominant kernel, cpu_baseline = the
reference's own CPU code (oracle/_ref) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per frame stream (before CUDA initialises)
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "Mpoints/s patch-gen+image-formation, longdress_vox10, 1/2/4/8 GPU"
HANDOFF = (2, 4, 5, 15, 16)  # what the video codec receives: OM video, geometry D0/D1 luma, padded attribute T0/T1 as 8-bit YUV 4:2:0


def pinned(shape, dtype):
    """numpy array over page-locked host memory (torch owns the allocation): H2D / D2H copies run at PCIe speed, no staging"""
    import torch
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    t = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    a = t.numpy()[:n].view(dtype).reshape(shape)
    _PINNED.append(t)
    return a


_PINNED = []


def make_frames(count, scale, seed=0, distinct=8, pin=False):
    import synth
    base = []
    for f in range(min(count, distinct)):
        xyz, rgb = synth.figure(scale=scale, seed=seed, frame=f)
        xyz, rgb = np.ascontiguousarray(xyz), np.ascontiguousarray(rgb)
        if pin:
            px, pc = pinned(xyz.shape, xyz.dtype), pinned(rgb.shape, rgb.dtype)
            px[...], pc[...] = xyz, rgb
            xyz, rgb = px, pc
        base.append((xyz, rgb))
    return [base[f % len(base)] for f in range(count)]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        if self.proc:
            self.proc.terminate()

        def num(s):
            return s.replace(".", "", 1).isdigit()

        sm = [float(r[0]) for r in self.rows if r and num(r[0])]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and num(r[1])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = {nme for r in self.rows for i, nme in enumerate(names) if len(r) > 2 + i and r[2 + i].lower().startswith("active")}
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    """cores this process may run on (the box's cgroup / affinity mask, not the machine's socket count)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# algorithmic bytes per launch of the single-kernel spans, N = points of the frame
ALGO_BYTES = {
    "knn16": lambda n: 6 * n + 64 * n,
    "normals": lambda n: 64 * n + 6 * n + 24 * n,
    "orient_walk": lambda n: 64 * n + 24 * n + 24 * n,
}


def config(args, npts, frames_per_rank, world):
    return {"workload": "synthetic longdress-like figure(scale=%.3f), 10-bit, %.2f Mpts/frame, %d frames/step/GPU (%d distinct), CTC all-intra r3 "
                        "(occupancyPrecision 4, I=%d refine iterations)" % (args.scale, npts / 1e6, frames_per_rank, min(frames_per_rank, 8), args.iterations),
            "stages": "a1-a26: kd-tree, k-NN16, PCA normals, spanning-tree orientation, initial + grid-refined segmentation, patch segmentation, "
                      "packing, occupancy/geometry images + dilation, generatePointCloud, colour transfer, attribute images, push-pull padding",
            "excluded": "ply load, videoEncoder.compress x3, post-processing, bitstream (as in BASELINE.md §4)",
            "frames_in_flight": frames_per_rank * max(1, min(args.gofs_in_flight, args.steps)), "gofs_in_flight": max(1, min(args.gofs_in_flight, args.steps)), "host_cores": host_cores(), "parallelism": "frames of a GOF sharded over %d GPU(s)" % world,
            "l2": "per-frame working set (>400 MB) and fresh uploads every step exceed the 126 MB L2",
            "host_buffers": "pinned (inputs and the frames handed to the video codec)",
            "handoff": "occupancy video + geometry D0/D1 luma + attribute T0/T1 converted to 8-bit YUV 4:2:0 on the device (the conversion the reference does inside compress())", "scratch_sets": args.scratch_sets}


def run_reference(args, frames, prm):
    """--impl reference: the reference's own CPU implementation (oracle/_ref, TBB off) on the host cores, one frame per process."""
    import multiprocessing as mp
    cores = max(1, min(len(frames), host_cores(), args.ref_frames))
    work = frames[:cores]
    ctx = mp.get_context("fork")

    def one(i):
        import bindings
        bindings.Reference().encode_gof([work[i]], prm)

    def step():
        ps = [ctx.Process(target=one, args=(i,)) for i in range(cores)]
        t0 = time.perf_counter()
        for p in ps:
            p.start()
        for p in ps:
            p.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup_ref):
        step()
    times = [step() for _ in range(args.steps_ref)]
    pts = sum(len(f[0]) for f in work)
    sec = float(np.mean(times))
    val = pts / sec / 1e6
    cfg = config(args, float(np.mean([len(f[0]) for f in work])), cores, 1)
    cfg["note"] = "reference TMC2 v24.0 compiled from /root/reference (ENABLE_TBB off, CTC --nbThread=1), one frame per host process"
    return {"metric": METRIC, "value": val, "unit": "Mpoints/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps_ref,
            "warmup": args.warmup_ref, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16/f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": "Mpoints/s", "cores": cores, "kind": "reference",
                             "sample": "%d frame(s) of the workload per step, one per host core (%d timed step(s) after %d warm-up)" % (cores, args.steps_ref, args.warmup_ref)},
            "e2e": {"value": val, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="frames per step and GPU (a GOF is 32 frames)")
    ap.add_argument("--scale", type=float, default=0.626, help="figure scale; 0.626 gives ~0.83 Mpts/frame like longdress_vox10")
    ap.add_argument("--iterations", type=int, default=50, help="iterationCountRefineSegmentation (longdress cfg: 50)")
    ap.add_argument("--ref-frames", type=int, default=32, help="frames per step of the reference arm (one host process per frame, up to the core count)")
    ap.add_argument("--gofs-in-flight", type=int, default=8, help="GOFs processed concurrently (each on its own context); steps are independent GOFs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timeline", default=None, help="write the last GOF's per-frame stage spans (name, start ms, ms) to this JSON file")
    ap.add_argument("--scratch-sets", type=int, default=24, help="scratch sets of the device pool = frames inside the data-parallel stage groups at once")
    args = ap.parse_args()
    # reference arm: every step is a bounded sample (one frame per host core, ~15 s); a few steps keep the run within minutes
    args.steps_ref, args.warmup_ref = max(1, min(args.steps, 3)), min(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import bindings

    if args.impl == "reference":
        if rank == 0:
            frames = make_frames(min(args.frames, args.ref_frames), args.scale, seed=0)
            weight = tuple(bindings.Reference().weight_normal(frames[0][0], 11))
            print(json.dumps(run_reference(args, frames, bindings.ctc_seg_params(bits=10, iterations=args.iterations, weight=weight))))
        return

    import torch
    dist = None
    real_stdout = os.dup(1)
    os.dup2(2, 1)   # NCCL / library chatter must not precede the JSON line on stdout: everything but the result goes to stderr
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
    frames = make_frames(args.frames, args.scale, seed=rank, pin=True)
    lanes = max(1, min(args.gofs_in_flight, args.steps))          # GOFs in flight: one library context (streams + buffers) each
    prods = [bindings.Product(local) for _ in range(lanes)]
    prod = prods[0]
    prod.lib.pccb200_set_scratch_sets(local, args.scratch_sets)
    # axis weights come from frame 0 of the GOF (rank 0's first frame); every rank needs the same three doubles
    w = torch.tensor(prod.weight_normal(frames[0][0], 11), dtype=torch.float64)
    if dist is not None:
        wd = w.cuda()
        dist.broadcast(wd, 0)
        w = wd.cpu()
    prm = bindings.ctc_seg_params(bits=10, iterations=args.iterations, weight=tuple(float(x) for x in w))
    for p in prods:
        p.profile(True)
    total_pts = sum(len(f[0]) for f in frames)
    outbuf, outlock = dict(), threading.Lock()   # ONE set of pinned hand-off buffers: the lanes take turns copying out (45 ms per GOF)

    def phase_a(lane):
        """a1..a13 of one GOF: segmentation (incl. the orientation walk) + packing"""
        t0 = time.perf_counter()
        g = bindings.ProductGof(prods[lane], frames, prm, 4)
        return g, (t0, time.perf_counter())

    def phase_b(lane, g, t0, W, H):
        """a16..a26 + hand-off: images, reconstruction, colour, attribute images; D2H of every frame the codec would receive"""
        p = prods[lane]
        t0, ta = t0
        g.resume(W, H, 0)
        t1 = time.perf_counter()
        nbytes = 0
        with outlock:
            for f in range(len(frames)):
                for what in HANDOFF:
                    cnt = g.count(f, what)
                    if (f, what) not in outbuf or outbuf[(f, what)].size != cnt:
                        outbuf[(f, what)] = pinned((cnt,), bindings.GOF_DTYPES[what])
                    g.fetch(f, what, outbuf[(f, what)])
                    nbytes += outbuf[(f, what)].nbytes
        t2 = time.perf_counter()
        spans = p.profile_read()
        g.free()
        host_phases.append((ta - t0, t1 - ta, t2 - t1, time.perf_counter() - t0))
        gof_log.append((lane, t0, ta, t1, t2, time.perf_counter()))
        return t1 - t0, t2 - t0, spans, nbytes

    gof_log = []       # per GOF: lane, absolute times of start / packed / resumed / fetched / freed
    host_phases = []   # per GOF: seconds in encode_gof(stop_after=1), all-reduce + resume, fetch, whole lane cycle

    def in_threads(fn, count):
        out = [None] * count
        ths = [threading.Thread(target=lambda i=i: out.__setitem__(i, fn(i))) for i in range(count)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return out

    stagger = [1.0]   # seconds between the starts of the lanes' first GOFs (re-estimated from the warm-up)

    def run_steps(count):
        """`count` independent GOFs. One lane: strictly one after the other. Several lanes: a software pipeline - every lane (own
        library context: streams + buffers) takes the next GOF as soon as it has finished its previous one, the lanes start
        staggered, so that one GOF's orientation walks (32 resident warps, GPU otherwise idle) overlap the data-parallel stages of
        the others. Per GOF: phase A, ONE all-reduce(MAX) of the canvas size (multi-GPU only, issued in GOF order on every rank),
        phase B."""
        results = [None] * count
        lock = threading.Condition()
        state = {"next": 0, "turn": 0}

        def worker(lane):
            torch.cuda.set_device(local)   # (the current device is per thread)
            if lane:
                time.sleep(lane * stagger[0])
            while True:
                with lock:
                    g = state["next"]
                    state["next"] += 1
                if g >= count:
                    return
                gof, t0 = phase_a(lane)
                W, H = gof.dims(0)[:2]
                if dist is not None:  # the one collective: common canvas size of the GOF across the ranks holding its frames
                    with lock:
                        while state["turn"] != g:
                            lock.wait()
                    wh = torch.tensor([W, H], device="cuda", dtype=torch.int64)
                    dist.all_reduce(wh, op=dist.ReduceOp.MAX)
                    W, H = (int(x) for x in wh.cpu())
                    with lock:
                        state["turn"] += 1
                        lock.notify_all()
                results[g] = phase_b(lane, gof, t0, W, H)

        in_threads(worker, min(lanes, count))
        return results

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 2 * lanes)   # every context (lane) allocates its buffers in its first GOF; its second one runs warm
    wres = run_steps(warm)
    if lanes > 1:   # steady-state spacing of GOF starts: latency of one (warm) GOF / lanes
        stagger[0] = float(np.min([e for _, e, _, _ in wres])) / lanes
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    del host_phases[:]
    del gof_log[:]
    t_begin = time.perf_counter()
    res = run_steps(args.steps)
    barrier()
    wall_total = time.perf_counter() - t_begin
    clocks = sampler.finish()
    last_spans, d2h = res[-1][2], res[-1][3]
    if args.timeline and rank == 0:
        with open(args.timeline, "w") as fh:
            json.dump({"host_phase_a_s": res[-1][0], "host_total_s": res[-1][1], "spans": [[n, float(st), float(ms)] for n, ms, st in last_spans]}, fh)
    # device window per GOF (first compute span to last span of any of its frame streams); GOFs overlap, so the job's device time
    # is bounded by the wall clock of the timed region: report the smaller of the two views consistently as wall-based
    dev_each = []
    for c, e, spans, nb in res:
        comp = [(st, st + ms) for nme, ms, st in spans if st >= 0 and nme not in ("h2d", "d2h_patches")]
        dev_each.append((max(b for a, b in comp) - min(a for a, b in comp)) / 1e3 if comp else c)
    e2e_t = wall_total
    dev_t = wall_total - (sum(e - c for c, e, _, _ in res) / max(1, lanes))   # wall minus the hand-off copies of one lane
    if lanes == 1:
        dev_t = float(np.sum(dev_each))
    if dist is not None:
        t = torch.tensor([e2e_t, dev_t, wall_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t, dev_t, wall_total = (float(x) for x in t)
    if rank != 0:
        return
    pts_all = total_pts * world * args.steps
    agg = {}
    for nme, ms, st in last_spans:
        agg.setdefault(nme, []).append(ms)
    mean_ms = {k: float(np.mean(v)) for k, v in agg.items()}
    dom = max((k for k in mean_ms if k in ALGO_BYTES), key=lambda k: mean_ms[k])
    peak, how = measured_peak()
    npts = float(np.mean([len(f[0]) for f in frames]))
    per_launch = args.frames if dom == "orient_walk" else 1   # the walks of all frames of a GOF share one launch
    algo_bytes = ALGO_BYTES[dom](npts) * per_launch
    ach = algo_bytes / (mean_ms[dom] * 1e-3) / 1e9
    h2d = sum(f[0].nbytes + f[1].nbytes for f in frames)
    out = {
        "metric": METRIC, "value": pts_all / dev_t / 1e6, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": wall_total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16/f64", "data": "synthetic", "config": config(args, npts, args.frames, world),
        "e2e": {"value": pts_all / e2e_t / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": None,
        "stage_ms_per_frame": {k: round(v, 3) for k, v in sorted(mean_ms.items(), key=lambda kv: -kv[1])},
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "peak_source": how, "algorithmic_bytes_per_launch": algo_bytes,
                     "note": "sequential spanning-tree walk, one warp per frame, all frames of a GOF in one launch: latency-bound by construction"},
        "clocks": clocks,
        "gof_device_window_ms": [round(x * 1e3, 1) for x in dev_each],
        "gof_log_ms": [[g[0]] + [round((x - t_begin) * 1e3) for x in g[1:]] for g in sorted(gof_log, key=lambda g: g[1])],
        "host_ms_per_gof": dict(zip(("segment_and_pack", "resume", "fetch", "cycle"), (round(float(np.median(c)) * 1e3, 1) for c in zip(*host_phases)))),
        "gpu_mem_used_gb": round((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30, 1),
    }
    counts_path = os.path.join(ROOT, "profiles", "launch_counts.json")
    if os.path.exists(counts_path):
        with open(counts_path) as f:
            lc = json.load(f)
        out["gpu_launches"] = int(lc.get("launches_per_frame", 0) * args.frames * args.steps)
        out["gpu_launches_source"] = lc.get("source")
        if lc.get("traffic_bytes_per_frame", {}).get(dom):   # ncu --set full of the kernel on ONE frame; a launch covers per_launch frames
            out["roofline"]["traffic"] = int(lc["traffic_bytes_per_frame"][dom] * per_launch)
            out["roofline"]["traffic_source"] = lc.get("traffic_source")
    if not args.no_cpu_baseline and os.path.exists(bindings.REF_SO):
        ref = bindings.Reference()
        t0 = time.perf_counter()
        ref.encode_gof([frames[0]], prm)
        sec = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": len(frames[0][0]) / sec / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "reference",
                               "sample": "1 frame of the workload (%.2f Mpts) through the reference's own stages, single thread (CTC --nbThread=1)" % (len(frames[0][0]) / 1e6)}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
