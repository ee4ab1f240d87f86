"""Host-side protocol of the frame-sharded GOF: what the ranks exchange, nothing else.

All-intra: the only cross-frame coupling is the common canvas size of a GOF (PCCEncoder::resizeGeometryVideo,
PccLibEncoder/source/PCCEncoder.cpp:5546-5591: the maximum over the frames, >= the minimum image size). `CanvasExchange` reduces
it with MAX over the ranks that hold the GOF's frames:

  * the sizes are host integers, so they are staged in ONE pre-allocated pinned buffer and reduced on a dedicated process group
    (NCCL: its own communicator and stream, created high-priority) from a dedicated comm thread - no allocation, no `.cpu()`, no
    default-stream work, nothing queued behind a frame's kernels;
  * several GOFs share one collective: it covers a window of up to `batch` consecutive GOFs starting at the oldest one that is
    not reduced yet, and carries a "not posted here yet" flag per slot (reduced with MAX like the sizes), so every rank learns
    from the result itself which slots are final and where the next window starts - no negotiation, no fixed batch boundaries,
    and a rank only ever waits for GOFs up to the one it asks for (a rank joins a collective once ITS oldest open slot is posted);
  * nobody waits for the collective before image formation: a rank goes ahead with its LOCAL canvas size (the GOF-wide maximum
    equals it unless another rank's frames needed more rows than the minimum image height) and checks the reduced size before
    the frames are handed to the codec; on a mismatch it re-forms the GOF on the larger canvas (`pccb200_gof_resume` accepts
    that on a finished GOF). Results are identical to the lock-step protocol; the collective is off the critical path.
"""
import threading


class CanvasExchange:
    def __init__(self, dist, total_gofs, batch=4, group=None, device=None):
        import torch
        self.torch, self.dist, self.total, self.batch, self.group = torch, dist, int(total_gofs), max(1, int(batch)), group
        self.cuda = device is not None
        self.lock = threading.Condition()
        self.local = {}      # gof -> (W, H) posted by this rank
        self.result = {}     # gof -> (W, H) reduced over the ranks
        self.error = None
        self.collectives = 0
        self.seconds = 0.0   # wall time spent inside the collectives (comm thread)
        if self.cuda:
            self.device = torch.device("cuda", device)
            self.stream = torch.cuda.Stream(device=self.device, priority=-1)
            self.host = torch.zeros(self.batch * 3, dtype=torch.int64).pin_memory()
            self.dev = torch.zeros(self.batch * 3, dtype=torch.int64, device=self.device)
            self.done = torch.cuda.Event(blocking=True)   # the comm thread sleeps, it does not spin on a core a frame thread needs
        else:
            self.host = torch.zeros(self.batch * 3, dtype=torch.int64)
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    # ---- lanes
    def post(self, gof, width, height):
        with self.lock:
            self.local[gof] = (int(width), int(height))
            self.lock.notify_all()

    def wait(self, gof):
        with self.lock:
            while gof not in self.result and self.error is None:
                self.lock.wait()
            if self.error is not None:
                raise RuntimeError("canvas exchange failed: %r" % (self.error,))
            return self.result[gof]

    def close(self):
        self.thread.join()
        if self.error is not None:
            raise RuntimeError("canvas exchange failed: %r" % (self.error,))

    # ---- comm thread: windows in GOF order; all ranks derive the same windows from the reduced results
    def _run(self):
        import time
        torch, dist = self.torch, self.dist
        try:
            if self.cuda:
                torch.cuda.set_device(self.device)
            first = 0
            while first < self.total:
                gofs = list(range(first, min(first + self.batch, self.total)))
                with self.lock:
                    while first not in self.local:
                        self.lock.wait()
                    for i, g in enumerate(gofs):
                        wh = self.local.get(g)
                        self.host[3 * i], self.host[3 * i + 1], self.host[3 * i + 2] = (wh[0], wh[1], 0) if wh else (0, 0, 1)
                    for i in range(len(gofs), self.batch):
                        self.host[3 * i], self.host[3 * i + 1], self.host[3 * i + 2] = 0, 0, 0
                t0 = time.perf_counter()
                if self.cuda:
                    with torch.cuda.stream(self.stream):
                        self.dev.copy_(self.host, non_blocking=True)
                        dist.all_reduce(self.dev, op=dist.ReduceOp.MAX, group=self.group)
                        self.host.copy_(self.dev, non_blocking=True)
                        self.done.record(self.stream)
                    self.done.synchronize()
                else:
                    dist.all_reduce(self.host, op=dist.ReduceOp.MAX, group=self.group)
                self.seconds += time.perf_counter() - t0
                self.collectives += 1
                with self.lock:
                    for i, g in enumerate(gofs):   # final up to the first slot some rank had not posted yet
                        if int(self.host[3 * i + 2]) != 0:
                            break
                        self.result[g] = (int(self.host[3 * i]), int(self.host[3 * i + 1]))
                        first = g + 1
                    self.lock.notify_all()
        except Exception as e:   # noqa: BLE001  (a lane blocked in wait() must see the failure)
            with self.lock:
                self.error = e
                self.lock.notify_all()


# ---------------------------------------------------------------------------------------------------- random access (a15)
# Every frame is packed against the previous one (PCCEncoder::placeSegments, PccLibEncoder/source/PCCEncoder.cpp:4778-4805) and the
# global patch allocation iterates over the whole GOF (:6838-6970): a sharded GOF needs the patch records (metadata + 16-pixel block
# occupancies, KBs per frame) of ALL its frames on every rank once - ONE all-gather; the packing itself is deterministic and is
# replicated (pccb200_gof_pack_ra), so no broadcast of the result is needed.
def gather_patch_records(dist, local, total_frames, patch_dtype, group=None, device=None):
    """local: [(global frame index, patch records (structured array), block occupancies (uint8))] of this rank's frames.
    Returns [(records, occupancies)] for frames 0..total_frames-1 on every rank. Variable sizes: one all-gather of the byte
    counts, one of the padded payloads."""
    import numpy as np
    import torch
    dev = torch.device("cuda", device) if device is not None else torch.device("cpu")
    world = dist.get_world_size(group)
    head = np.zeros((len(local), 3), np.int64)
    blobs = []
    for i, (f, patches, occ) in enumerate(local):
        pb = np.ascontiguousarray(patches, patch_dtype).tobytes()
        ob = np.ascontiguousarray(occ, np.uint8).tobytes()
        head[i] = (f, len(pb), len(ob))
        blobs += [pb, ob]
    payload = np.frombuffer(np.array([len(local)], np.int64).tobytes() + head.tobytes() + b"".join(blobs), np.uint8)
    size = torch.tensor([payload.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    most = max(int(s) for s in sizes)
    mine = torch.zeros(most, dtype=torch.uint8, device=dev)
    mine[:payload.size] = torch.from_numpy(payload.copy()).to(dev)
    parts = [torch.zeros(most, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = [None] * total_frames
    for r in range(world):
        raw = parts[r][:int(sizes[r])].cpu().numpy().tobytes()
        n = int(np.frombuffer(raw[:8], np.int64)[0])
        hd = np.frombuffer(raw[8:8 + 24 * n], np.int64).reshape(n, 3)
        at = 8 + 24 * n
        for f, pb, ob in hd:
            patches = np.frombuffer(raw[at:at + pb], patch_dtype).copy()
            occ = np.frombuffer(raw[at + pb:at + pb + ob], np.uint8).copy()
            out[int(f)] = (patches, occ)
            at += int(pb + ob)
    if any(o is None for o in out):
        raise RuntimeError("patch records of some frames of the GOF were not gathered")
    return out


class RecordExchange:
    """The all-gathers of a run of sharded random-access GOFs, issued in GOF order by a comm thread on its own communicator and
    stream (like CanvasExchange; here the lanes DO wait for the result: the packing of a GOF needs every rank's records)."""

    def __init__(self, dist, total_gofs, frames_per_gof, patch_dtype, group=None, device=None):
        import torch
        self.torch, self.dist, self.total, self.frames, self.dtype, self.group = torch, dist, int(total_gofs), int(frames_per_gof), patch_dtype, group
        self.device = device
        self.lock = threading.Condition()
        self.local, self.result, self.error = {}, {}, None
        self.collectives, self.seconds = 0, 0.0
        self.stream = torch.cuda.Stream(device=torch.device("cuda", device), priority=-1) if device is not None else None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def post(self, gof, local_records):
        with self.lock:
            self.local[gof] = local_records
            self.lock.notify_all()

    def wait(self, gof):
        with self.lock:
            while gof not in self.result and self.error is None:
                self.lock.wait()
            if self.error is not None:
                raise RuntimeError("record exchange failed: %r" % (self.error,))
            return self.result.pop(gof)

    def close(self):
        self.thread.join()
        if self.error is not None:
            raise RuntimeError("record exchange failed: %r" % (self.error,))

    def _run(self):
        import contextlib
        import time
        try:
            if self.device is not None:
                self.torch.cuda.set_device(self.device)
            for g in range(self.total):
                with self.lock:
                    while g not in self.local:
                        self.lock.wait()
                    mine = self.local.pop(g)
                t0 = time.perf_counter()
                with (self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()):
                    got = gather_patch_records(self.dist, mine, self.frames, self.dtype, group=self.group, device=self.device)
                self.seconds += time.perf_counter() - t0
                self.collectives += 2
                with self.lock:
                    self.result[g] = got
                    self.lock.notify_all()
        except Exception as e:   # noqa: BLE001
            with self.lock:
                self.error = e
                self.lock.notify_all()
