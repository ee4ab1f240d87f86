import sys, time
sys.path.insert(0,'/root/repo/tests')
import numpy as np, bindings, synth
orc=bindings.Oracle()
scale=float(sys.argv[1]) if len(sys.argv)>1 else 0.3
xyz,rgb=synth.figure(scale=scale, seed=0, frame=0)
N=len(xyz); print('N',N)
idx,_=orc.knn(xyz,xyz,16)
nrm=orc.normals(xyz, idx, False)
src=np.repeat(np.arange(N,dtype=np.int64),16); dst=idx.ravel().astype(np.int64)
keep=(dst!=src)&(dst<N)
src,dst=src[keep],dst[keep]
w=np.abs((nrm[src]*nrm[dst]).sum(1))
E=len(src); print('edges',E)
# mutual flag
key=src*N+dst; rkey=dst*N+src
ks=np.sort(key); pos=np.searchsorted(ks,rkey); pos[pos>=E]=E-1
mutual=ks[pos]==rkey
print('mutual fraction',mutual.mean())
label=np.arange(N,dtype=np.int64)
def find_all(label):
    while True:
        l2=label[label]
        if np.array_equal(l2,label): return label
        label=l2
for rnd in range(60):
    ls,ld=label[src],label[dst]
    ext=ls!=ld
    if not ext.any(): break
    mx=np.full(N,-1.0)
    np.maximum.at(mx,ls[ext],w[ext]); np.maximum.at(mx,ld[ext],w[ext])
    # contractible: external, mutual, dominant at both supernodes
    c=ext&mutual&(w>=mx[ls])&(w>=mx[ld])
    if not c.any(): break
    a,b=ls[c],ld[c]
    lo,hi=np.minimum(a,b),np.maximum(a,b)
    # each supernode has at most one dominant edge => pairs are a matching; union by pointing hi->lo
    label[hi]=lo
    label=find_all(label)
    ns=len(np.unique(label))
    print('round',rnd,'contracted pairs',int(c.sum())//2,'supernodes',ns, 'ratio %.3f'%(ns/N))
ls,ld=label[src],label[dst]; ext=ls!=ld
mx=np.full(N,-1.0); np.maximum.at(mx,ls[ext],w[ext]); np.maximum.at(mx,ld[ext],w[ext])
# why blocked: the dominant external edge of each supernode is one-way?
dom=ext&((w>=mx[ls])|(w>=mx[ld]))
print('blocked supernodes', len(np.unique(label)), 'dominant edges mutual fraction', mutual[dom].mean())
sizes=np.bincount(np.unique(label,return_inverse=True)[1]); print('max supernode',sizes.max(),'mean',sizes.mean())
