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
