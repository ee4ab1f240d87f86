"""N>1 path on CPU (gloo, world_size 2): frames of one GOF are sharded over ranks; the only exchange is the all-reduce(MAX) of the
canvas size between packing and image formation (and the broadcast of frame 0's axis weights). The compute stand-in is the
oracle (no GPU here); the protocol, the collective and the equivalence 'sharded == unsharded' are what is tested."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _frames():
    import synth
    # the tall canvas comes from frame 2 only (many patches), so rank 0 must learn the height from rank 1
    return [synth.double_sheet(n_side=40, seed=1), synth.sphere(radius=18, center=60, seed=2), synth.figure(scale=0.16, seed=3, frame=0),
            synth.specks(seed=4)]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    import bindings
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = bindings.Oracle()
    frames = _frames()
    mine = frames[rank::world]                                    # frame f -> rank f mod world
    w = torch.zeros(3, dtype=torch.float64)
    if rank == 0:
        w = torch.tensor(orc.weight_normal(frames[0][0], 11))     # axis weights of frame 0 of the GOF
    dist.broadcast(w, 0)
    prm = bindings.ctc_seg_params(bits=10, iterations=6, weight=tuple(float(x) for x in w))
    packed = orc.encode_gof(mine, prm, stop_after=1)              # a1..a13 on the local shard
    wh = torch.tensor([packed[0].width, packed[0].height], dtype=torch.int64)
    dist.all_reduce(wh, op=dist.ReduceOp.MAX)                     # the one cross-frame reduction (resizeGeometryVideo)
    full = orc.encode_gof(mine, prm, canvas=(int(wh[0]), int(wh[1])))
    np.save(os.path.join(out_dir, "canvas_%d.npy" % rank), wh.numpy())
    for i, fr in enumerate(full):
        np.savez(os.path.join(out_dir, "rank%d_frame%d.npz" % (rank, i)), **{str(k): v for k, v in fr.data.items()},
                 patches=fr.patches.patches, local_h=np.array([packed[0].height]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gof_equals_unsharded(tmp_path, oracle):
    import bindings
    world, port = 2, 29731 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    frames = _frames()
    prm = bindings.ctc_seg_params(bits=10, iterations=6, weight=oracle.weight_normal(frames[0][0], 11))
    whole = oracle.encode_gof(frames, prm)
    canvases = [np.load(tmp_path / ("canvas_%d.npy" % r)) for r in range(world)]
    assert all(tuple(c) == (whole[0].width, whole[0].height) for c in canvases), "all-reduced canvas != GOF canvas"
    local_heights = set()
    for f, fr in enumerate(whole):
        rank, i = f % world, f // world
        got = np.load(tmp_path / ("rank%d_frame%d.npz" % (rank, i)))
        local_heights.add(int(got["local_h"][0]))
        assert np.array_equal(got["patches"], fr.patches.patches)
        for k, v in fr.data.items():
            assert np.array_equal(got[str(k)], v), "frame %d product %s differs between sharded and unsharded runs" % (f, bindings.GOF_NAMES[k])
    assert len(local_heights) > 1 or whole[0].height == 1280, "test should exercise a shard whose local canvas is smaller"
