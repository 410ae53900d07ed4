"""Multi-GPU host plumbing: one process per GPU, torch.distributed for rendezvous only.

BA shards POINTS (with their observations, CSR-contiguous) across ranks; cameras and the focal length are replicated
(SURVEY.md section 8e).  The only data-path exchange is the sum of the reduced camera system inside libsfmb200.so
(NCCL, csrc/comm.cu); this module only splits problems and hands the NCCL unique id around.
"""
import numpy as np


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [begin, end) of n_items for `rank` of `world` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_ba_problem(p, rank, world):
    """Strong-scaling split of a flattened BA problem (dict as produced by synth.make_ba_problem): this rank's points and
    their observations, ALL cameras."""
    b, e = shard_bounds(p["np"], rank, world)
    o0, o1 = int(p["pt_off"][b]), int(p["pt_off"][e])
    out = dict(p)
    out.update(np=e - b, nobs=o1 - o0, pts=p["pts"][b:e], obs_xy=p["obs_xy"][o0:o1], obs_cam=p["obs_cam"][o0:o1],
               pt_off=(p["pt_off"][b:e + 1] - o0).astype(np.int32), point_range=(b, e))
    if "obs_pt" in p:
        out["obs_pt"] = (p["obs_pt"][o0:o1] - b).astype(np.int32)
    if "pts_true" in p:
        out["pts_true"] = p["pts_true"][b:e]
    return out


def shard_pairs(pairs, rank, world):
    """Image pairs are independent (the reference shards them over threads, SfM.cpp:166-206): round-robin over ranks."""
    return [pr for i, pr in enumerate(pairs) if i % world == rank]


def extract_features_sharded(images, rank, world, extract, dist_module=None):
    """SfM::extractFeatures (SfM.cpp:141-154) over several GPUs: images are independent, so rank r extracts the contiguous chunk
    shard_bounds(len(images), r, world) with `extract(list_of_images) -> list of per-image results` (stages.extractAllFeatures on its own
    GPU) and the per-image results are gathered on the host in image order on every rank -- no collective on the data path
    ("chunk split only")."""
    b, e = shard_bounds(len(images), rank, world)
    mine = extract(list(images[b:e])) if e > b else []
    if world == 1:
        return mine
    dist = dist_module
    if dist is None:
        import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    return [f for part in parts for f in part]


def init_comm(ctx, dist_module=None):
    """Create the library's NCCL communicator for the current torch.distributed world (no-op for world size 1)."""
    from . import capi
    dist = dist_module
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    return world


def attach_peers(prob, dist_module=None):
    """Switch a BAProblem's cross-rank sums from NCCL to peer-memory kernels: all-gather the CUDA-IPC handles of the
    exchange buffers (rank order) and attach them.  Collective: every rank must call it after creating its problem."""
    dist = dist_module
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return False
    handles = [None] * dist.get_world_size()
    dist.all_gather_object(handles, prob.ipc_handle())
    prob.ipc_attach(handles)
    dist.barrier()
    return True
