"""How predictable is the orientation walk? Simulates the reference's greedy walk (max |n.n| edge leaving the visited set) on a
synthetic frame and reports, for every visited point, how many visits ago the edge that reached it was queued ("age"): age 1 =
the point is a neighbour of the point visited just before (the kernel's popNew case), small ages = still in the 32-entry hot
set, larger = the bit-tree. Usage: python tools/walk_age_stats.py [scale]"""
import heapq
import sys
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import bindings
import synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
xyz, _ = synth.figure(scale=scale, seed=0, frame=0)
orc = bindings.Oracle()
idx, _ = orc.knn(xyz, xyz, 16)
nrm = orc.normals(xyz, idx, False)
N = len(xyz)
w = np.abs(np.einsum("ij,ikj->ik", nrm, nrm[idx.astype(np.int64) % N]))
visited = np.zeros(N, bool)
ages = []
step = 0
for seed in range(N):
    if visited[seed]:
        continue
    visited[seed] = True
    heap = []
    def push(u):
        for s in range(16):
            v = int(idx[u, s])
            if v < N and not visited[v]:
                heapq.heappush(heap, (-w[u, s], -u, -v, step))
    push(seed)
    while heap:
        negw, nu, nv, t = heapq.heappop(heap)
        v = -nv
        if visited[v]:
            continue
        visited[v] = True
        step += 1
        ages.append(step - t)
        push(v)
ages = np.array(ages)
print("points", N, "visits", len(ages))
for a in (1, 2, 4, 8, 16, 32, 64, 256, 1024):
    print("age <= %4d : %.1f %%" % (a, 100.0 * (ages <= a).mean()))
