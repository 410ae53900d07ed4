"""Multi-GPU equivalence (needs >= 2 GPUs; skipped otherwise): a BA problem sharded by points over 2 ranks, with the
reduced camera system summed by NCCL inside libsfmb200.so, gives the 1-GPU result to fp64 round-off."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, out_dir, exchange):
    import torch.distributed as dist
    from sfm_toy_library_b200 import capi, synth
    from sfm_toy_library_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # host plumbing only (unique-id hand-off)
    ctx = capi.Context(rank)
    assert sdist.init_comm(ctx, dist) == world and ctx.comm_size == world
    p = synth.make_ba_problem(n_cams=16, n_pts=3001, obs_per_pt=6, seed=8)
    sh = sdist.shard_ba_problem(p, rank, world)
    prob = ctx.ba_problem(sh["cams"], sh["pts"], sh["focal"], sh["obs_xy"], sh["obs_cam"], sh["pt_off"])
    if exchange == "peer":
        assert sdist.attach_peers(prob, dist)                       # CUDA-IPC mapped buffers, sums by NVLink loads
    red = prob.reduced_system(1e4)
    if exchange == "oneshot":
        # the drop-in call (sfmb200_ba_solve): attaches its peers itself (IPC handles over one ncclAllGather), no host-side
        # exchange, no barrier before the problem is destroyed (the run ends with a peer hand-shake)
        prob.close()
        cams, pts, f, s = ctx.ba_solve(sh["cams"], sh["pts"], sh["focal"], sh["obs_xy"], sh["obs_cam"], sh["pt_off"])
        cams2, pts2, f2, s2 = ctx.ba_solve(sh["cams"], sh["pts"], sh["focal"], sh["obs_xy"], sh["obs_cam"], sh["pt_off"])   # cached mappings
        assert s2["num_iterations"] == s["num_iterations"] and np.array_equal(cams, cams2) and np.array_equal(pts, pts2)
    else:
        s = prob.run()
        cams, pts, f = prob.download()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), S=red["S"], rhs=red["rhs"], cost=red["cost"], cams=cams, pts=pts, f=f,
             iters=s["num_iterations"], term=s["termination_type"], final=s["final_cost"], b=sh["point_range"][0], e=sh["point_range"][1])
    prob.close(); ctx.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("exchange", ["nccl", "peer", "oneshot"])
def test_two_gpu_solve_equals_single_gpu(tmp_path, exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from sfm_toy_library_b200 import capi, synth
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), exchange), nprocs=2, join=True)
    r = [np.load(os.path.join(tmp_path, f"r{i}.npz")) for i in range(2)]
    p = synth.make_ba_problem(n_cams=16, n_pts=3001, obs_per_pt=6, seed=8)
    ctx = capi.Context(0)
    prob = ctx.ba_problem(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"])
    red = prob.reduced_system(1e4); s = prob.run(); cams, pts, f = prob.download()
    for x in r:                                                       # every rank holds the full, summed reduced system
        np.testing.assert_allclose(x["S"], red["S"], rtol=0, atol=1e-11 * np.abs(red["S"]).max())
        np.testing.assert_allclose(x["rhs"], red["rhs"], rtol=0, atol=1e-10 * np.abs(red["rhs"]).max())
        assert abs(float(x["cost"]) - red["cost"]) < 1e-12 * red["cost"]
        assert int(x["iters"]) == s["num_iterations"] and int(x["term"]) == s["termination_type"] == capi.CONVERGENCE
        assert abs(float(x["final"]) - s["final_cost"]) < 1e-10 * s["final_cost"]
        np.testing.assert_allclose(x["cams"], cams, rtol=0, atol=1e-8)
        np.testing.assert_allclose(x["pts"], pts[int(x["b"]):int(x["e"])], rtol=0, atol=1e-8)
    np.testing.assert_array_equal(r[0]["cams"], r[1]["cams"])         # identical decisions and cameras on every rank
    prob.close(); ctx.close()
