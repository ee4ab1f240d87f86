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


# ---------------------------------------------------------------------------------------------------------------------
# the protocol bench.py runs (mpeg-pcc-tmc2_b200/sharding.py): batched all-reduce from a comm thread, image formation on the LOCAL
# canvas size first, re-formed on the reduced size when another rank needed more rows
def _exchange_worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "mpeg-pcc-tmc2_b200"))
    import bindings
    from sharding import CanvasExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    group = dist.new_group(backend="gloo")
    orc = bindings.Oracle()
    frames = _frames()
    w = torch.tensor(orc.weight_normal(frames[0][0], 11)) if rank == 0 else torch.zeros(3, dtype=torch.float64)
    dist.broadcast(w, 0)
    prm = bindings.ctc_seg_params(bits=10, iterations=6, weight=tuple(float(x) for x in w))
    # five GOFs: in GOF 0 and 3 the ranks hold different frames (rank 1's need the taller canvas), in the others the same ones
    gofs = [frames[rank::world], frames[:2], frames[:2], frames[rank::world], frames[1:2]]
    ex = CanvasExchange(dist, len(gofs), batch=2, group=group, device=None)
    reformed = []
    import threading

    def lane(which):   # two lanes per rank, GOFs taken alternately in increasing order (as bench.py does: a lane never waits for a GOF it will post later)
        for g in which:
            packed = orc.encode_gof(gofs[g], prm, stop_after=1)
            W, H = packed[0].width, packed[0].height
            ex.post(g, W, H)
            full = orc.encode_gof(gofs[g], prm, canvas=(W, H))            # ahead on the local size
            Wg, Hg = ex.wait(g)
            if (Wg, Hg) != (W, H):
                reformed.append(g)
                full = orc.encode_gof(gofs[g], prm, canvas=(Wg, Hg))
            np.savez(os.path.join(out_dir, "x_rank%d_gof%d.npz" % (rank, g)), canvas=np.array([Wg, Hg]),
                     **{"f%d_%d" % (i, k): v for i, fr in enumerate(full) for k, v in fr.data.items()})
    ths = [threading.Thread(target=lane, args=([0, 2, 4],)), threading.Thread(target=lane, args=([1, 3],))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    ex.close()
    np.save(os.path.join(out_dir, "x_stats_%d.npy" % rank), np.array([ex.collectives, len(reformed)]))
    dist.barrier()
    dist.destroy_process_group()


def test_canvas_exchange_protocol(tmp_path, oracle):
    import bindings
    world, port = 2, 29331 + (os.getpid() % 200)
    mp.spawn(_exchange_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    frames = _frames()
    prm = bindings.ctc_seg_params(bits=10, iterations=6, weight=oracle.weight_normal(frames[0][0], 11))
    whole = oracle.encode_gof(frames, prm)                       # GOF 0 / 3 unsharded: frames 0..3
    stats = [np.load(tmp_path / ("x_stats_%d.npy" % r)) for r in range(world)]
    assert stats[0][0] == stats[1][0] and 3 <= int(stats[0][0]) <= 5, "5 GOFs in windows of up to 2: 3..5 collectives, the same on every rank"
    assert sum(int(s[1]) for s in stats) >= 1 or whole[0].height == 1280, "a rank with the smaller local canvas must re-form"
    for g in (0, 3):
        for rank in range(world):
            got = np.load(tmp_path / ("x_rank%d_gof%d.npz" % (rank, g)))
            assert tuple(got["canvas"]) == (whole[0].width, whole[0].height)
            for i, f in enumerate(range(rank, len(frames), world)):
                for k, v in whole[f].data.items():
                    assert np.array_equal(got["f%d_%d" % (i, k)], v), "GOF %d frame %d product %s" % (g, f, bindings.GOF_NAMES[k])


# ---------------------------------------------------------------------------------------------------------------------
# random access: the one exchange of a sharded GOF (mpeg-pcc-tmc2_b200/sharding.py gather_patch_records) - every rank must end up with
# the records of all frames, identical to what a single process sees; feeding them to the packer then gives the unsharded packing
# (the product's pccb200_gof_pack_ra needs a GPU: tests/test_ra_pack.py; here the oracle's packer is the stand-in)
def _gather_worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "mpeg-pcc-tmc2_b200"))
    import bindings
    from sharding import gather_patch_records
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = bindings.Oracle()
    frames = _frames() + [(np.zeros((0, 3), np.int16), np.zeros((0, 3), np.uint8))]      # incl. an empty frame
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=(1.0, 1.0, 1.0))
    local = []
    for f in range(rank, len(frames), world):
        seg = orc.segment_frame_patches(frames[f][0], frames[f][1], prm) if len(frames[f][0]) else None
        patches = seg.patches if seg is not None else np.zeros(0, bindings.PATCH_DTYPE)
        occ = seg.occ if seg is not None else np.zeros(0, np.uint8)
        local.append((f, patches, occ))
    got = gather_patch_records(dist, local, len(frames), bindings.PATCH_DTYPE)
    np.savez(os.path.join(out_dir, "g_rank%d.npz" % rank), **{"p%d" % f: got[f][0] for f in range(len(frames))},
             **{"o%d" % f: got[f][1] for f in range(len(frames))})
    dist.barrier()
    dist.destroy_process_group()


def test_gather_patch_records(tmp_path, oracle):
    import bindings
    world, port = 2, 29531 + (os.getpid() % 200)
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    frames = _frames() + [(np.zeros((0, 3), np.int16), np.zeros((0, 3), np.uint8))]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=(1.0, 1.0, 1.0))
    ranks = [np.load(tmp_path / ("g_rank%d.npz" % r)) for r in range(world)]
    for f in range(len(frames)):
        if len(frames[f][0]):
            seg = oracle.segment_frame_patches(frames[f][0], frames[f][1], prm)
            want_p, want_o = seg.patches, seg.occ
        else:
            want_p, want_o = np.zeros(0, bindings.PATCH_DTYPE), np.zeros(0, np.uint8)
        for r in range(world):
            assert np.array_equal(ranks[r]["p%d" % f], want_p) and np.array_equal(ranks[r]["o%d" % f], want_o), "frame %d on rank %d" % (f, r)
