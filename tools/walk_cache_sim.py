"""What would a shared-memory front of best[] buy the orientation walk? Simulates the reference's greedy walk on a synthetic frame
(same walk as tools/walk_age_stats.py) and replays the kernel's best[] gathers - per visited point the 16 neighbours, addressed by
kd-tree position as orient.cu does - against (a) a one-bit-per-point "touched" map (an entry never written holds "not on the
frontier": no memory access needed) and (b) a direct-mapped cache of recently accessed entries, for several cache sizes.
Reports the fraction of gathered entries and of STEPS (a step waits for its slowest lane) that would still go to L2 / DRAM.
Usage: python tools/walk_cache_sim.py [scale]"""
import heapq
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np  # noqa: E402
import bindings  # noqa: E402
import synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
xyz, _ = synth.figure(scale=scale, seed=0, frame=0)
orc = bindings.Oracle()
idx, _ = orc.knn(xyz, xyz, 16)
nrm = orc.normals(xyz, idx, False)
vind = orc.vind(xyz)                     # tree order -> caller index
N = len(xyz)
pos = np.empty(N, np.int64)
pos[vind.astype(np.int64)] = np.arange(N)
w = np.abs(np.einsum("ij,ikj->ik", nrm, nrm[idx.astype(np.int64) % N]))
visited = np.zeros(N, bool)
order = []
for seed in range(N):
    if visited[seed]:
        continue
    visited[seed] = True
    order.append(seed)
    heap = []

    def push(u):
        for s in range(16):
            v = int(idx[u, s])
            if v < N and not visited[v]:
                heapq.heappush(heap, (-w[u, s], -u, -v))
    push(seed)
    while heap:
        _, nu, nv = heapq.heappop(heap)
        v = -nv
        if visited[v]:
            continue
        visited[v] = True
        order.append(v)
        push(v)
order = np.array(order)
print("points %d, visits %d" % (N, len(order)))
touched = np.zeros(N, bool)
sizes = [2048, 4096, 8192, 16384, 32768]
tags = {c: np.full(c, -1, np.int64) for c in sizes}
entry_first = entry_total = 0
entry_miss = {c: 0 for c in sizes}
step_any_first = 0
step_miss = {c: 0 for c in sizes}
step_miss_with_map = {c: 0 for c in sizes}
for u in order:
    nb = idx[u].astype(np.int64)
    nb = nb[nb < N]
    p = pos[nb]
    first = ~touched[p]
    entry_total += len(p)
    entry_first += int(first.sum())
    step_any_first += bool(first.any())
    for c in sizes:
        t = tags[c]
        slot = p % c
        hit = t[slot] == p
        entry_miss[c] += int((~hit & ~first).sum())
        step_miss[c] += bool((~hit).any())                   # cache alone: a first touch is a miss too
        step_miss_with_map[c] += bool((~hit & ~first).any())  # touched map answers the first touches
        t[slot] = p
    touched[p] = True
    touched[pos[u]] = True
print("gathered entries: %d, never written before (answered by the touched map): %.1f %%; steps with such an entry: %.1f %%"
      % (entry_total, 100.0 * entry_first / entry_total, 100.0 * step_any_first / len(order)))
for c in sizes:
    print("cache %6d entries (%3d KB): entries still missing %.1f %%; steps that wait for memory: cache alone %.1f %%, with the touched map %.1f %%"
          % (c, c * 8 // 1024, 100.0 * entry_miss[c] / entry_total, 100.0 * step_miss[c] / len(order), 100.0 * step_miss_with_map[c] / len(order)))
