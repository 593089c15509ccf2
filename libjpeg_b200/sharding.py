"""Multi-GPU host logic of the decode path: frames are independent, so rank r simply owns a contiguous shard of the
batch; the ONLY collective is the broadcast of the shared Huffman/quantisation table blob from rank 0 (NCCL on GPUs,
gloo in the CPU tests).  Everything here is host code; the blob is built by b200jpg_build_tables (no GPU needed)."""
import ctypes

import numpy as np

from . import native
from .native import lib


def shard_range(total, rank, world):
    """Contiguous, balanced shard [begin, end) of `total` frames for `rank` of `world`."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def build_tables(codestream, scan=0):
    """Device-independent table blob (uint8 numpy array) of one scan of a codestream."""
    data = np.frombuffer(bytes(codestream), dtype=np.uint8)
    size = lib.b200jpg_build_tables(data.ctypes.data, data.size, scan, None, 0)
    if size == 0:
        code, msg = native.last_error(None)
        raise native.NativeError(code, msg)
    out = np.zeros(size, dtype=np.uint8)
    lib.b200jpg_build_tables(data.ctypes.data, data.size, scan, out.ctypes.data, size)
    return out


def broadcast_tables(blob, dist, device=None):
    """Rank 0's blob to every rank (one collective). Returns the received blob as a numpy array and whether this
    rank's own blob was identical (frames with different tables cannot share the broadcast and must keep their own)."""
    import torch
    mine = torch.from_numpy(np.ascontiguousarray(blob))
    size = torch.tensor([mine.numel()], dtype=torch.int64)
    if device is not None:
        mine, size = mine.to(device), size.to(device)
    dist.broadcast(size, src=0)
    recv = mine.clone() if int(size.item()) == mine.numel() else torch.zeros(int(size.item()), dtype=torch.uint8, device=mine.device)
    dist.broadcast(recv, src=0)
    same = recv.numel() == mine.numel() and bool(torch.equal(recv, mine))
    return recv.cpu().numpy(), same
