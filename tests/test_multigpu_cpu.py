"""world_size-2 gloo tests (CPU) of the N>1 host logic: partitions and the sort-last key composite."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenes


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_partitions():
    from taichi_three_b200 import multigpu as M
    for n, w in ((64, 8), (64, 3), (5, 8), (0, 2)):
        views = sorted(sum((M.view_partition(n, r, w) for r in range(w)), []))
        assert views == list(range(n))
        ranges = [M.face_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1


def test_pack_unpack_roundtrip():
    from taichi_three_b200 import multigpu as M
    rng = np.random.default_rng(0)
    depth = rng.integers(-2**30, 2**30, (17, 9)).astype(np.int32)
    occup = rng.integers(-1, 1000, (17, 9)).astype(np.int32)
    keys = M.pack_keys(depth, occup, face_base=12345)
    d, o = M.unpack_keys(keys, face_base=12345, nfaces=1000)
    assert np.array_equal(d.numpy(), depth) and np.array_equal(o.numpy(), occup)
    # ordering: depth first (signed), then face id; "no face" sorts before any face at equal depth
    a = M.pack_keys(np.int32([5]), np.int32([7]))
    b = M.pack_keys(np.int32([5]), np.int32([3]))
    c = M.pack_keys(np.int32([-4]), np.int32([900]))
    e = M.pack_keys(np.int32([5]), np.int32([-1]))
    assert c < b < a and e < b
    assert int(M.pack_keys(np.int32([2**30]), np.int32([-1]))) == M.CLEAR_KEY


def _worker(rank, world, port, tri, W, H, W2V, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import oracle as O
    from taichi_three_b200 import multigpu as M
    lo, hi = M.face_range(len(tri), rank, world)
    occup, depth, _, _ = O.render_occup(tri[lo:hi], W2V, W, H)
    keys = M.pack_keys(depth, occup, face_base=lo)
    M.composite_min(keys)
    d, o = M.unpack_keys(keys)  # global ids
    mine = M.unpack_keys(keys, face_base=lo, nfaces=hi - lo)[1]
    owned = torch.tensor([(mine >= 0).sum()], dtype=torch.int64)
    dist.all_reduce(owned)
    if rank == 0:
        np.savez(out, depth=d.numpy(), occup=o.numpy(), owned=owned.numpy())
    dist.destroy_process_group()


def test_sort_last_composite_equals_single_pass_gloo(tmp_path, O):
    """Two ranks rasterise disjoint face ranges with global ids, MIN-all-reduce the packed keys:
    bit-identical to one pass over all faces (ties included), every covered pixel owned by one rank."""
    W, H = 96, 64
    view, proj = scenes.default_camera(W / H)
    W2V = (proj @ view).astype(np.float32)
    tri = scenes.soup(3000, W, H, s=0.06, seed=8)
    tri[1500:1700] = tri[100:300]  # exact inter-rank depth ties: the lower global id must win
    ref_occup, ref_depth, tie, _ = O.render_occup(tri, W2V, W, H)
    out = str(tmp_path / 'r0.npz')
    mp.spawn(_worker, args=(2, _free_port(), tri, W, H, W2V, out), nprocs=2, join=True)
    r = np.load(out)
    assert np.array_equal(r['depth'], ref_depth)
    assert np.array_equal(r['occup'], ref_occup)
    assert int(r['owned'][0]) == int((ref_occup >= 0).sum())
    assert not np.isin(ref_occup, np.arange(1500, 1700)).any()


def test_strip_ranges_tile_the_screen():
    from taichi_three_b200 import multigpu as M
    for npix, w in ((7680 * 4320, 8), (1920 * 1080, 8), (1000, 3), (256, 4), (3840 * 2160, 2)):
        r = [M.strip_range(npix, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == npix
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        assert all(lo % 256 == 0 for lo, hi in r if lo < npix)
        for share in (0.0, 0.5):  # the image root takes a smaller strip (sort-last over peer memory)
            r = [M.strip_range(npix, k, w, root=0, root_share=share) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == npix and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert all(lo % 256 == 0 for lo, hi in r if lo < npix)
            if share == 0.0:
                assert r[0] == (0, 0)
    assert M.sort_last_strip(7680 * 4320, 0, 8)[:2] == (0, 0) and M.sort_last_strip(7680 * 4320, 0, 2)[2] == 1.0
