"""Host-side helpers for the one-process-per-GPU layout: read sharding and NCCL id distribution.

Reads shard naturally: rank r scans the contiguous read-id range [bounds[r], bounds[r+1]) (balanced by bases);
the k-mer table is replicated. The exchanges that follow the scan live in libhga_b200.so (hga_comm_init)."""
import numpy as np


def shard_bounds(read_lengths, world):
    """Contiguous read ranges with (nearly) equal base counts. Returns world+1 read indices."""
    lens = np.asarray(read_lengths, dtype=np.int64)
    csum = np.cumsum(lens)
    total = int(csum[-1]) if lens.shape[0] else 0
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, total * r // world)))
    bounds.append(int(lens.shape[0]))
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def broadcast_unique_id(dist, rank, make_id):
    """rank 0 creates the 128-byte NCCL unique id (make_id()), everyone receives it through torch.distributed."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise ValueError("NCCL unique id must be 128 bytes")
    return bytes(uid)
