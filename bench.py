#!/usr/bin/env python
"""bench.py — Mpoints/s of the V-PCC patch-generation + image-formation hot path (SURVEY.md §8a rows a1–a26).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F] [--scale S]
                  [--condition ai|ra] [--bits 10|11] [--ply-dir DIR]

Workloads (BASELINE.json configs): default = configs[1] (longdress-like, 10 bit, 0.83 Mpts/frame, CTC all-intra r3, I=50);
`--condition ra` = configs[2] (random-access r5: occupancyPrecision 2, global patch allocation); `--bits 11` = configs[4]
(~2.9 Mpts/frame, 2560-wide canvas, I=20; use --frames 8 --gofs-in-flight 8 --scratch-sets 8: a frame keeps ~1.6 GB while in flight);
`--ply-dir DIR` replaces the synthetic figure by real frames (e.g. longdress_vox10_1051..1082.ply) read by pccb200_ply_read.

One "step" = one GOF (--frames F frames per GPU) of synthetic longdress-like frames (tests/synth.py `figure`, 10-bit, ≈0.83 M points
per frame at the default scale) pushed through the whole hot path: generateSegments, placeSegments, occupancy / geometry
image formation, generatePointCloud, colour transfer, attribute image formation and padding (PCCEncoder.cpp:103-424,
the three videoEncoder.compress calls excluded, occupancy/geometry video treated as lossless).  No dataset ships with
the reference, so the data is synthetic and says so.

Steps are independent GOFs; --gofs-in-flight of them are processed concurrently (one library context each, started staggered):
the orientation walk of a frame is a 1-2 s single-warp latency chain, so the GPU is kept busy by the data-parallel stages of
the other GOFs in flight. value / e2e are whole-job throughputs over the timed region (K GOFs, barrier to barrier).

Multi-GPU (torchrun, one process per GPU): the frames of every GOF are sharded over the ranks (rank r takes F frames of
an N*F-frame GOF: weak scaling); the one cross-frame coupling of the all-intra path — the common canvas size — is an
NCCL all-reduce(MAX) issued by a comm thread on its own communicator / high-priority stream, several GOFs per collective, and
taken off the critical path: image formation goes ahead on the local size and the reduced size is checked before the GOF is
handed off (mpeg-pcc-tmc2_b200/sharding.py). Random access (--condition ra) shards the same way; every frame is packed against the
previous one there, so the ranks all-gather the patch records (KBs of host data per frame: a gloo group) once per GOF and every rank
runs the deterministic packing (pccb200_gof_pack_ra).  Timing: barrier + synchronize on both sides, max over ranks.

JSON keys follow the driver contract.  value = throughput over the DEVICE window (first compute span to last span of any
frame stream, CUDA events; input already uploaded), e2e = wall clock through the C ABI with host buffers including the
H2D of the clouds and the D2H of every frame handed to the video codec (per-lane pinned buffers), roofline = the stage span with the
longest duration plus a per-stage table and the whole-path figure, cpu_baseline = the reference's own CPU code (oracle/_ref) on a
bounded sample (one frame), against whose products frame 0 of this run is compared ("parity_checked" / "parity_ok").
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
HANDOFF = (2, 17, 18, 15, 16)  # what the video codec receives: OM video, geometry D0/D1 luma as bytes, padded attribute T0/T1 as 8-bit YUV 4:2:0


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


def read_ply_frames(directory, count):
    """real frames: every .ply of the directory in name order through the product's own reader (PCCPointSet3::read semantics)"""
    import ctypes as C
    import bindings
    lib = C.CDLL(bindings.PRODUCT_SO)
    lib.pccb200_ply_read.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    names = sorted(n for n in os.listdir(directory) if n.lower().endswith(".ply"))[:count]
    if not names:
        raise SystemExit("no .ply files in %s" % directory)
    out = []
    for nme in names:
        path = os.path.join(directory, nme).encode()
        n, col = C.c_size_t(0), C.c_int(0)
        if lib.pccb200_ply_read(path, None, None, 0, C.byref(n), C.byref(col)) != 0:
            raise SystemExit("cannot read %s" % nme)
        xyz, rgb = np.zeros((n.value, 3), np.int16), np.zeros((n.value, 3), np.uint8)
        if lib.pccb200_ply_read(path, xyz.ctypes.data_as(C.c_void_p), rgb.ctypes.data_as(C.c_void_p), n.value, C.byref(n), C.byref(col)) != 0:
            raise SystemExit("cannot read %s" % nme)
        out.append((xyz, rgb))
    return out


def make_frames(count, args, seed=0, distinct=8, pin=False):
    import synth
    if args.ply_dir:
        base = read_ply_frames(args.ply_dir, count)
    else:
        base = [synth.figure(scale=args.scale, seed=seed, bits=args.bits, frame=f) for f in range(min(count, distinct))]
    out = []
    for xyz, rgb in base:
        xyz, rgb = np.ascontiguousarray(xyz), np.ascontiguousarray(rgb)
        if pin:
            px, pc = pinned(xyz.shape, xyz.dtype), pinned(rgb.shape, rgb.dtype)
            px[...], pc[...] = xyz, rgb
            xyz, rgb = px, pc
        out.append((xyz, rgb))
    return [out[f % len(out)] for f in range(count)]


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


# Algorithmic bytes per frame of every stage span (every input streamed once, every output written once, a gather
# counted as one pass over the gathered array). q: N source points, R reconstructed points, Q canvas pixels, I refine iterations,
# prec occupancy precision. The refinement and patch-segmentation terms use §8d's worked-example model for the quantities the
# kernels do not report (V = N/14 voxels, a quarter of them active with 64 adjacency entries, T = 4 outer iterations over
# N, N/4, N/16, N/64 raw points, 60 smoothing passes over the push-pull pyramid): "model" in the roofline note.
ALGO_BYTES = {
    "kdtree_build": lambda q: 8 * q["N"] + 4 * q["N"] + 8 * q["N"] + 16 * (2 * q["N"] / 10),
    "knn16": lambda q: 6 * q["N"] + 64 * q["N"],
    "normals": lambda q: 64 * q["N"] + 6 * q["N"] + 24 * q["N"],
    "orient": lambda q: 64 * q["N"] + 24 * q["N"] + 128 * q["N"],
    "orient_walk": lambda q: 64 * q["N"] + 24 * q["N"] + 24 * q["N"],
    "initial_seg": lambda q: 24 * q["N"] + q["N"],
    "refine": lambda q: (7 * q["N"] + 12 * q["N"] / 14) + q["I"] * (12 * 0.25 * (q["N"] / 14) * 64 + 12 * q["N"] / 14 + 26 * 0.25 * q["N"]),
    "patches": lambda q: 69 * 1.33 * q["N"] + 12 * q["N"] + 4 * q["R"] + 4 * (6 * q["N"] + 6 * q["R"] + 8 * q["N"]),
    "images": lambda q: q["Q"] / q["prec"] ** 2 * 1.5 + 2 * q["Q"] * 1.5 + 2 * 2 * q["Q"] * 2,
    "reconstruct": lambda q: q["Q"] / q["prec"] ** 2 + 4 * q["Q"] + 22 * q["R"],
    "color_transfer": lambda q: (6 * q["R"] + 6 * q["N"] + 32 * q["R"]) + (6 * q["N"] + 6 * q["R"] + 4 * q["N"]) + 3 * q["N"] + 3 * q["R"],
    "attribute_images": lambda q: 2 * q["Q"] * 3 + 2 * (4.0 / 3.0) * q["Q"] * 3 * (2 + 2 * 60),
}


def workload_name(args, npts):
    cond = "all-intra r3 (occupancyPrecision 4)" if args.condition == "ai" else "random-access r5 (occupancyPrecision 2, global patch allocation)"
    src = "real .ply frames from %s" % args.ply_dir if args.ply_dir else "synthetic longdress-like figure(scale=%.3f)" % args.scale
    return "%s, %d-bit, %.2f Mpts/frame, CTC %s, I=%d refine iterations" % (src, args.bits, npts / 1e6, cond, args.iterations)


def config(args, npts, frames_per_rank, world):
    lanes = max(1, min(args.gofs_in_flight, args.steps))
    sharded = args.condition == "ai"
    return {"workload": workload_name(args, npts) + ", %d frames/step/GPU (%d distinct)" % (frames_per_rank, min(frames_per_rank, 8)),
            "stages": "a1-a26: kd-tree, k-NN16, PCA normals, spanning-tree orientation, initial + grid-refined segmentation, patch segmentation, "
                      "packing, occupancy/geometry images + dilation, generatePointCloud, colour transfer, attribute images, push-pull padding",
            "excluded": "ply load, videoEncoder.compress x3, post-processing, bitstream (as in BASELINE.md §4)",
            "frames_in_flight": frames_per_rank * lanes, "gofs_in_flight": lanes, "host_cores": host_cores(),
            "parallelism": ("frames of a GOF sharded over %d GPU(s), canvas size reduced by NCCL all-reduce(MAX), up to %d GOFs per collective, off the critical path"
                            % (world, args.exchange_batch)) if sharded else
                           "frames of a GOF sharded over %d GPU(s) (frame f -> GPU f mod G), one all-gather of the patch records (host data, gloo) per GOF, packing replicated" % world,
            "l2": "per-frame working set (>400 MB) and fresh uploads every step exceed the 126 MB L2",
            "host_buffers": "pinned (inputs and the frames handed to the video codec)",
            "handoff": "occupancy video + geometry D0/D1 luma as bytes + attribute T0/T1 converted to 8-bit YUV 4:2:0 on the device (the conversion the reference does inside compress())",
            "scratch_sets": args.scratch_sets}


def run_reference(args, frames, prm, prec):
    """--impl reference: the reference's own CPU implementation (oracle/_ref, TBB off) on the host cores, one frame per process."""
    import multiprocessing as mp
    cores = max(1, min(len(frames), host_cores(), args.ref_frames))
    work = frames[:cores]
    ctx = mp.get_context("fork")

    def one(i):
        import bindings
        bindings.Reference().encode_gof([work[i]], prm, occupancy_precision=prec)

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
            "dtype": "int16/f64", "data": "real" if args.ply_dir else "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": "Mpoints/s", "cores": cores, "kind": "reference",
                             "sample": "%d frame(s) of the workload per step, one per host core (%d timed step(s) after %d warm-up)" % (cores, args.steps_ref, args.warmup_ref)},
            "e2e": {"value": val, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def build_params(args, weight):
    import bindings
    prm = bindings.ctc_seg_params(bits=args.bits, iterations=args.iterations, weight=weight)
    if args.condition == "ra":
        prm.global_patch_allocation = 1   # cfg/condition/ctc-random-access.cfg (constrainedPack stays at its default 1)
    return prm


def parity_check(prod, frame0, prm, prec, handoff_of_last_gof):
    """frame 0 of the workload through the CUDA path against the reference's own CPU run of the same frame (oracle/_ref), every
    product; without the compiled reference, against the reference-generated digests of tests/golden/gof_fullsize.json.
    Returns (cpu seconds or None, dict for the JSON line)."""
    import bindings
    got = prod.encode_gof([frame0], prm, occupancy_precision=prec)
    info = {"parity_checked": False}
    sec = None
    if os.path.exists(bindings.REF_SO):
        ref = bindings.Reference()
        t0 = time.perf_counter()
        want, _ = ref.encode_gof([frame0], prm, occupancy_precision=prec)
        sec = time.perf_counter() - t0
        bad = bindings.compare_gof(got, want)
        info = {"parity_checked": True, "parity_ok": bad == [], "parity_against": "oracle/_ref (the compiled reference) on frame 0 of the workload: "
                "patch list + depth/occupancy arenas + all 16 products (a13-a26, YUV hand-off), bit-exact", "parity_mismatches": bad[:6]}
    else:
        import hashlib
        path = os.path.join(ROOT, "tests", "golden", "gof_fullsize.json")
        if os.path.exists(path):
            with open(path) as f:
                gold = json.load(f)["cases"]
            key = {("ai", 10): "ai_r3", ("ra", 10): "ra_r5", ("ai", 11): "vox11"}.get((prm.global_patch_allocation and "ra" or "ai", prm.geometry_bitdepth_3d - 1))
            if key in gold and gold[key]["points"][0] == len(frame0[0]) and (got[0].width, got[0].height) == (gold[key]["frames"][0]["width"], gold[key]["frames"][0]["height"]) and key != "ra_r5":
                g0 = gold[key]["frames"][0]
                bad = [n for w, n in bindings.GOF_NAMES.items() if hashlib.sha256(np.ascontiguousarray(got[0].data[w]).tobytes()).hexdigest() != g0[n]]
                info = {"parity_checked": True, "parity_ok": bad == [], "parity_against": "tests/golden/gof_fullsize.json (sha256 of the reference's products, frame 0)", "parity_mismatches": bad[:6]}
    # the frames the TIMED pipeline handed off are the ones just checked (same frame, same canvas)
    if handoff_of_last_gof is not None and info.get("parity_checked"):
        def checked(what):   # the byte forms of the geometry planes are GEO0 / GEO1 of the checked frame, narrowed
            return got[0].data[what - 13].astype(np.uint8) if what in (17, 18) else got[0].data[what]
        same = all(np.array_equal(buf, checked(what)) for what, buf in handoff_of_last_gof.items() if buf.size == checked(what).size)
        info["timed_handoff_matches_checked_frame"] = bool(same)
    return sec, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="frames per step and GPU (a GOF is 32 frames)")
    ap.add_argument("--scale", type=float, default=None, help="figure scale; default 0.626 (10 bit: ~0.83 Mpts/frame like longdress_vox10) / 0.585 (11 bit: ~2.9 Mpts)")
    ap.add_argument("--bits", type=int, default=10, choices=[10, 11], help="geometry3dCoordinatesBitdepth (11: basketball_player_vox11, 2560-wide canvas)")
    ap.add_argument("--condition", default="ai", choices=["ai", "ra"], help="ai: CTC all-intra r3 (BASELINE configs[1]); ra: CTC random-access r5 (configs[2])")
    ap.add_argument("--iterations", type=int, default=None, help="iterationCountRefineSegmentation (longdress cfg: 50; basketball_player: 20)")
    ap.add_argument("--ply-dir", default=None, help="directory of .ply frames (sorted by name) to use instead of the synthetic figure; read by pccb200_ply_read")
    ap.add_argument("--ref-frames", type=int, default=32, help="frames per step of the reference arm (one host process per frame, up to the core count)")
    ap.add_argument("--gofs-in-flight", type=int, default=8, help="GOFs processed concurrently (each on its own context); steps are independent GOFs")
    ap.add_argument("--exchange-batch", type=int, default=4, help="GOFs whose canvas sizes share one all-reduce (multi-GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timeline", default=None, help="write the last GOF's per-frame stage spans (name, start ms, ms) to this JSON file")
    ap.add_argument("--scratch-sets", type=int, default=24, help="scratch sets of the device pool = frames inside the data-parallel stage groups at once")
    args = ap.parse_args()
    if args.scale is None:
        args.scale = 0.626 if args.bits == 10 else 0.585
    if args.iterations is None:
        args.iterations = 50 if args.bits == 10 else 20
    prec = 4 if args.condition == "ai" else 2
    # reference arm: every step is a bounded sample (one frame per host core, ~15 s); a few steps keep the run within minutes
    args.steps_ref, args.warmup_ref = max(1, min(args.steps, 3)), min(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import bindings

    if args.impl == "reference":
        if rank == 0:
            frames = make_frames(min(args.frames, args.ref_frames), args, seed=0)
            weight = tuple(bindings.Reference().weight_normal(frames[0][0], args.bits + 1))
            print(json.dumps(run_reference(args, frames, build_params(args, weight), prec)))
        return

    import torch
    dist = None
    real_stdout = os.dup(1)
    os.dup2(2, 1)   # NCCL / library chatter must not precede the JSON line on stdout: everything but the result goes to stderr
    exchange_group = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
        if args.condition == "ai":   # the canvas exchange gets its own NCCL communicator + high-priority stream (hidden behind image formation)
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            exchange_group = dist.new_group(backend="nccl", pg_options=opts)
        else:
            # Random access: the lanes WAIT for the exchanged patch records (the packing needs them), and the records are KBs of HOST
            # data (pccb200_gof_patches). Measured on 2 x B200 (profiles/r02j_bench_2gpu_ra_nccl.json): staged through the busy GPUs and
            # NCCL, one all-gather took 1.2 s (the collective kernels and copies queue behind the frames' kernels) and the ranks spent
            # 5.4 s per GOF waiting. Host data goes over a host transport: a gloo group beside the NCCL one.
            exchange_group = dist.new_group(backend="gloo")
    sharded = world > 1
    sharded_ra = sharded and args.condition == "ra"
    sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc2_b200"))
    from sharding import CanvasExchange, RecordExchange
    frames = make_frames(args.frames, args, seed=rank, pin=True)
    lanes = max(1, min(args.gofs_in_flight, args.steps))          # GOFs in flight: one library context (streams + buffers) each
    prods = [bindings.Product(local) for _ in range(lanes)]
    prod = prods[0]
    prod.lib.pccb200_set_scratch_sets(local, args.scratch_sets)
    # axis weights come from frame 0 of the GOF (rank 0's first frame); every rank needs the same three doubles
    w = torch.tensor(prod.weight_normal(frames[0][0], args.bits + 1), dtype=torch.float64)
    if sharded:
        wd = w.cuda()
        dist.broadcast(wd, 0)
        w = wd.cpu()
    prm = build_params(args, tuple(float(x) for x in w))
    for p in prods:
        p.profile(True)
    total_pts = sum(len(f[0]) for f in frames)
    outbuf = dict()   # pinned hand-off buffers, one set per lane (no lock: the lanes' copies overlap instead of queueing behind each other)

    def phase_a(lane):
        """a1..a13 of one GOF: segmentation (incl. the orientation walk) + packing (sharded random access: segmentation only)"""
        return bindings.ProductGof(prods[lane], frames, prm, prec, stop_after=5 if sharded_ra else 1)

    # sharded random access: local frame i of a rank is frame i * world + rank of the GOF (frame f -> GPU f mod G)
    total_frames = args.frames * world
    local_of = [(f // world if f % world == rank else -1) for f in range(total_frames)]

    def phase_b(lane, g, W, H):
        """a16..a26 + hand-off: images, reconstruction, colour, attribute images; D2H of every frame the codec would receive"""
        g.resume(W, H, 0)
        t1 = time.perf_counter()
        nbytes = 0
        for f in range(len(frames)):
            for what in HANDOFF:
                cnt = g.count(f, what)
                key = (lane, f, what)
                if key not in outbuf or outbuf[key].size != cnt:
                    outbuf[key] = pinned((cnt,), bindings.GOF_DTYPES.get(what) or bindings.GOF_EXTRA_DTYPES[what])
                g.fetch(f, what, outbuf[key])
                nbytes += outbuf[key].nbytes
        return t1, time.perf_counter(), nbytes

    gof_log = []       # per GOF: lane, absolute times of start / packed / resumed / fetched / verified / freed
    host_phases = []   # per GOF: seconds in segment_and_pack, image formation, fetch, wait for the reduced canvas size, whole cycle
    stats = {"reformed": 0}

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
        the others. Per GOF: phase A; the canvas size is posted to the exchange (multi-GPU: one NCCL all-reduce(MAX) per
        `--exchange-batch` GOFs, issued by the comm thread in GOF order); phase B goes ahead on the local size and the reduced size
        is checked before the GOF counts as done (re-formed on the larger canvas if another rank needed more rows)."""
        results = [None] * count
        lock = threading.Lock()
        state = {"next": 0}
        ex = CanvasExchange(dist, count, batch=args.exchange_batch, group=exchange_group, device=local) if sharded and not sharded_ra else None
        rx = RecordExchange(dist, count, total_frames, bindings.PATCH_DTYPE, group=exchange_group, device=None) if sharded_ra else None

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
                t0 = time.perf_counter()
                gof = phase_a(lane)
                tw = 0.0
                if rx is not None:   # ONE exchange of patch records (KBs per frame), then the deterministic packing on every rank
                    rx.post(g, [(i * world + rank,) + gof.patch_records(i) for i in range(len(frames))])
                    tx = time.perf_counter()
                    records = rx.wait(g)
                    tw = time.perf_counter() - tx
                    gof.pack_ra(records, local_of)
                ta = time.perf_counter()
                W, H = gof.dims(0)[:2]
                if ex is not None:
                    ex.post(g, W, H)
                t1, t2, nbytes = phase_b(lane, gof, W, H)
                tv = t2
                if ex is not None:
                    Wg, Hg = ex.wait(g)
                    tv = time.perf_counter()
                    if (Wg, Hg) != (W, H):   # another rank's frames needed a larger canvas: form this GOF again on it
                        stats["reformed"] += 1
                        _, _, nbytes = phase_b(lane, gof, Wg, Hg)
                spans = prods[lane].profile_read()
                gof.free()
                tf = time.perf_counter()
                host_phases.append((ta - t0, t1 - ta, t2 - t1, (tv - t2) + tw, tf - t0))
                gof_log.append((lane, t0, ta, t1, t2, tv, tf))
                results[g] = (t1 - t0, t2 - t0, spans, nbytes, lane)

        in_threads(worker, min(lanes, count))
        xs = None
        if ex is not None:
            ex.close()
            xs = (ex.collectives, ex.seconds)
        if rx is not None:
            rx.close()
            xs = (rx.collectives, rx.seconds)
        return results, xs

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # initialisation (untimed, not a warm-up step): every lane's context allocates its buffers while processing its first GOF
    ires, _ = run_steps(lanes)
    wres, _ = run_steps(args.warmup) if args.warmup > 0 else ([], None)
    if lanes > 1:   # steady-state spacing of GOF starts: latency of one (warm) GOF / lanes
        stagger[0] = float(np.min([r[1] for r in (wres or ires)])) / lanes
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    del host_phases[:]
    del gof_log[:]
    stats["reformed"] = 0
    t_begin = time.perf_counter()
    res, xstats = run_steps(args.steps)
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
    for c, e, spans, nb, _ in res:
        comp = [(st, st + ms) for nme, ms, st in spans if st >= 0 and nme not in ("h2d", "d2h_patches")]
        dev_each.append((max(b for a, b in comp) - min(a for a, b in comp)) / 1e3 if comp else c)
    e2e_t = wall_total
    dev_t = wall_total - (sum(r[1] - r[0] for r in res) / max(1, lanes))   # wall minus the hand-off copies of one lane
    if lanes == 1:
        dev_t = float(np.sum(dev_each))
    if dist is not None:
        t = torch.tensor([e2e_t, dev_t, wall_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t, dev_t, wall_total = (float(x) for x in t)
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pts_all = total_pts * world * args.steps
    agg = {}
    for nme, ms, st in last_spans:
        agg.setdefault(nme, []).append(ms)
    mean_ms = {k: float(np.mean(v)) for k, v in agg.items()}
    peak, how = measured_peak()
    npts = float(np.mean([len(f[0]) for f in frames]))
    # sizes of the last GOF for the algorithmic-bytes formulas
    g = bindings.ProductGof(prod, frames[:1], prm, prec)
    W, H = g.dims(0)[:2]
    g.resume(W, H, 0)
    R = g.dims(0)[2]
    g.free()
    prod.profile_read()
    q = {"N": npts, "R": float(R), "Q": float(W * H), "I": float(args.iterations), "prec": float(prec)}
    # a span's duration is per frame, except the walk: all frames of a GOF walk in ONE launch (one CTA each)
    per_launch = {k: (args.frames if k == "orient_walk" else 1) for k in ALGO_BYTES}
    table = {}
    for k, fn in ALGO_BYTES.items():
        if k in mean_ms and mean_ms[k] > 0:
            b = fn(q) * per_launch[k]
            table[k] = {"ms": round(mean_ms[k], 3), "algorithmic_bytes": int(b), "achieved_gbs": round(b / (mean_ms[k] * 1e-3) / 1e9, 3),
                        "frac": b / (mean_ms[k] * 1e-3) / 1e9 / peak}
    dom = max(table, key=lambda k: table[k]["ms"])
    algo_bytes = table[dom]["algorithmic_bytes"]
    ach = algo_bytes / (mean_ms[dom] * 1e-3) / 1e9
    path_bytes = sum(fn(q) for fn in ALGO_BYTES.values())                       # per frame, whole hot path
    path_gbs = path_bytes * (pts_all / npts) / dev_t / 1e9                        # frames processed / device time
    h2d = sum(f[0].nbytes + f[1].nbytes for f in frames)
    out = {
        "metric": METRIC, "value": pts_all / dev_t / 1e6, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "init_gofs": lanes, "ms_per_step": wall_total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16/f64", "data": "real" if args.ply_dir else "synthetic", "config": config(args, npts, args.frames, world),
        "e2e": {"value": pts_all / e2e_t / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": None,
        "stage_ms_per_frame": {k: round(v, 3) for k, v in sorted(mean_ms.items(), key=lambda kv: -kv[1])},
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "peak_source": how, "algorithmic_bytes_per_launch": algo_bytes,
                     "note": "dominant stage span of the last timed GOF (spans overlap the other GOFs in flight, so durations are under load); "
                             "refine / patches / attribute_images bytes use the SURVEY 8d model constants",
                     "per_stage": table,
                     "whole_path": {"algorithmic_bytes_per_frame": int(path_bytes), "achieved_gbs": round(path_gbs, 2), "frac": path_gbs / peak / max(1, world)}},
        "clocks": clocks,
        "gof_device_window_ms": [round(x * 1e3, 1) for x in dev_each],
        "gof_log_ms": [[g[0]] + [round((x - t_begin) * 1e3) for x in g[1:]] for g in sorted(gof_log, key=lambda g: g[1])],
        "host_ms_per_gof": dict(zip(("segment_and_pack", "form_images", "fetch", "canvas_wait", "cycle"), (round(float(np.median(c)) * 1e3, 1) for c in zip(*host_phases)))),
        "gpu_mem_used_gb": round((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30, 1),
    }
    if xstats is not None:
        out["exchange"] = {"kind": "patch-record all-gather (host data over gloo, 2 collectives per GOF)" if sharded_ra else "canvas-size all-reduce(MAX) over NCCL",
                           "collectives": xstats[0], "ms_per_collective": round(xstats[1] / max(1, xstats[0]) * 1e3, 2), "gofs_reformed": stats["reformed"]}
    counts_path = os.path.join(ROOT, "profiles", "launch_counts.json")
    if os.path.exists(counts_path):
        with open(counts_path) as f:
            lc = json.load(f)
        out["gpu_launches"] = int(lc.get("launches_per_frame", 0) * args.frames * args.steps)
        out["gpu_launches_source"] = lc.get("source")
        if lc.get("traffic_bytes_per_frame", {}).get(dom):   # ncu dram bytes of the stage's kernels on ONE frame; a launch covers per_launch frames
            out["roofline"]["traffic"] = int(lc["traffic_bytes_per_frame"][dom] * per_launch[dom])
            out["roofline"]["traffic_source"] = lc.get("traffic_source")
    if not args.no_cpu_baseline:
        last_lane = res[-1][4]
        last = {what: outbuf[(last_lane, 0, what)] for what in HANDOFF if (last_lane, 0, what) in outbuf} if args.frames > 0 else None
        sec, info = parity_check(prod, frames[0], prm, prec, last)
        out.update(info)
        if sec is not None:
            out["cpu_baseline"] = {"value": len(frames[0][0]) / sec / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "reference",
                                   "sample": "1 frame of the workload (%.2f Mpts) through the reference's own stages, single thread (CTC --nbThread=1)" % (len(frames[0][0]) / 1e6)}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
