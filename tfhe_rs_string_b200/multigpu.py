"""Multi-GPU plumbing for the KS+PBS path: one process per GPU, server key broadcast ONCE over NCCL,
independent ciphertext shards afterwards (the reference's analogue is rayon's par_iter over
independent ciphertexts, benches/core_crypto/pbs_bench.rs:517-531).  No data-path collective exists
because the path has no exchange step."""
import numpy as np


def shard_bounds(total, world, rank):
    """Contiguous split of `total` independent units over `world` ranks; the first total % world
    ranks take one extra unit.  Returns (begin, end)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("invalid world/rank")
    base, extra = divmod(int(total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def level_shards(level_sizes, world, rank):
    """Per dependency level of a program (ragged sizes), the [begin, end) range this rank bootstraps."""
    return [shard_bounds(n, world, rank) for n in level_sizes]


def arena_tensor(ptr, nbytes, device):
    """Zero-copy torch view of a device allocation owned by the library (the key arena)."""
    import torch

    class _Arena:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Arena(), device=device)


def broadcast_server_key(engine, dist, rank, device, src=0):
    """Rank `src` has loaded the keys (H2D + std->Fourier + byte-limb conversion on its GPU); every
    other rank receives the finished arena (Fourier BSK || KSK || column sums || KSK limbs) in one
    broadcast and adopts it."""
    import torch
    ptr, nbytes = engine.key_arena()
    t = arena_tensor(ptr, nbytes, device)
    dist.broadcast(t, src=src)
    torch.cuda.synchronize()
    if rank != src:
        engine.keys_adopt()
    return nbytes


def broadcast_bytes(buf, dist, src=0):
    """Backend-agnostic form of the same step (used by the gloo CPU test): in-place broadcast of a
    uint8 tensor."""
    dist.broadcast(buf, src=src)
    return buf


def gather_counts(total, world):
    return np.array([shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)])
