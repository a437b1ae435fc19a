"""Camera-rig sharding across the GPUs of one node.

Cameras are independent -- one CPSNWhere_Tracker2D per camera with no shared state
(psn_where/PSNWhere.h:21, psn_where/PSNWhere.cpp:257-266) -- so camera c lives on rank c mod world_size, each rank
owns the pyramid rings of its cameras, and nothing crosses GPUs on the data path.  The only exchange is the
host-side gather of the per-camera 2D results that feed CPSNWhere_Associator3D::Run on rank 0
(psn_where/PSNWhere.cpp:269): a few hundred PSN_Rect per camera.
"""
import numpy as np


def cameras_of_rank(n_cameras, world_size, rank):
    """Global camera ids owned by `rank` (round robin)."""
    return list(range(rank, n_cameras, world_size))


def rank_of_camera(cam, world_size):
    return cam % world_size


def local_index(cam, world_size):
    """Index of global camera `cam` inside its rank's psnw context."""
    return cam // world_size


def gather_results(dist, local, n_cameras, dst=0):
    """Gather {global camera id: (ids int32 [k], boxes float64 [k,4])} from every rank to `dst`.

    `dist` is torch.distributed (gloo or nccl process group already initialised) or None for a single process.
    Returns the merged dict ordered by camera id on `dst`, None elsewhere."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        merged = dict(local)
    else:
        payload = {int(c): (np.asarray(i, np.int32), np.asarray(b, np.float64).reshape(-1, 4)) for c, (i, b) in local.items()}
        out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
        dist.gather_object(payload, out, dst=dst)
        if dist.get_rank() != dst:
            return None
        merged = {}
        for part in out:
            for c, v in part.items():
                if c in merged:
                    raise ValueError("camera %d reported by two ranks" % c)
                merged[c] = v
    missing = [c for c in range(n_cameras) if c not in merged]
    if missing:
        raise ValueError("no result for cameras %s" % missing)
    return {c: merged[c] for c in sorted(merged)}
