"""N > 1 host logic on CPU: two gloo ranks shard a BA problem by points, each computes its partial reduced camera system
with the oracle, the all-reduced sum equals the unsharded system -- the additivity the GPU path relies on when it
all-reduces S | rhs | gradient | cost."""
import os
import socket

import numpy as np
import pytest

from sfm_toy_library_b200 import dist as sdist
from sfm_toy_library_b200 import synth


def test_shard_bounds_cover_without_overlap():
    for n in (0, 1, 7, 200_000):
        for world in (1, 2, 3, 8):
            spans = [sdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_shard_pairs_round_robin():
    pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)]
    got = sorted(sum((sdist.shard_pairs(pairs, r, 4) for r in range(4)), []))
    assert got == sorted(pairs)


def test_sharded_feature_extraction_single_rank_is_the_plain_call():
    imgs = [np.full((4, 4), i, np.uint8) for i in range(5)]
    out = sdist.extract_features_sharded(imgs, 0, 1, lambda chunk: [int(im[0, 0]) * 10 for im in chunk])
    assert out == [0, 10, 20, 30, 40]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = synth.make_ba_problem(n_cams=5, n_pts=101, obs_per_pt=3, seed=4)
    sh = sdist.shard_ba_problem(p, rank, world)
    # undamped, unscaled system: additive over points
    part = oracle.ba_reduced_system(sh["cams"], sh["pts"], sh["focal"], sh["obs_xy"], sh["obs_cam"], sh["pt_off"], radius=1e300,
                                    jacobi_scaling=0, min_diag=0.0)
    n = 6 * p["nc"]
    buf = torch.from_numpy(np.concatenate([part["S"].ravel(), part["rhs"], part["grad"][:n], part["grad"][-1:], [part["cost"]]]))
    dist.all_reduce(buf)                                      # what ncclAllReduce does on the device
    full = oracle.ba_reduced_system(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"], radius=1e300,
                                    jacobi_scaling=0, min_diag=0.0)
    ref = np.concatenate([full["S"].ravel(), full["rhs"], full["grad"][:n], full["grad"][-1:], [full["cost"]]])
    np.testing.assert_allclose(buf.numpy(), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    # the point gradient is the shard's own slice
    b, e = sh["point_range"]
    np.testing.assert_allclose(part["grad"][n:n + 3 * (e - b)], full["grad"][n + 3 * b:n + 3 * e], rtol=1e-10, atol=1e-12)
    # init_comm is a no-op without the CUDA library's communicator when world == 1 and needs the GPU otherwise:
    # here we only check that the unique-id hand-off would reach every rank
    obj = [b"x" * 128 if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert obj[0] == b"x" * 128
    # ORB extraction shards images (row f-3): every rank extracts its chunk, the host gathers the per-image results in image order;
    # the extractor here is the ORB oracle (CPU), the GPU stage is bit-identical to it
    from oracle import orb_oracle
    rs = np.random.RandomState(7)
    imgs = [rs.randint(0, 256, (96, 128), dtype=np.uint8) for _ in range(3)]
    calls = []

    def extract(chunk):
        calls.append(len(chunk))
        return [orb_oracle.detect_and_compute(im, 200) for im in chunk]
    feats = sdist.extract_features_sharded(imgs, rank, world, extract)
    assert calls == [2 if rank == 0 else 1] and len(feats) == 3
    for im, (k, d) in zip(imgs, feats):
        ko, do = orb_oracle.detect_and_compute(im, 200)
        assert np.array_equal(k, ko) and np.array_equal(d, do)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_reduced_system_is_additive(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok{r}")) for r in range(world))
