"""Image-parallel inversion across the GPUs of one box.

Every op on the generator path is batch-independent, so images shard across ranks with no
data-path collective.  The only exchange is ONE all-gather per step of a packed per-image
record  [w+ (9*256) | decoder latent (n_latent*512) | K metric scalars]  — what the
reference's single-process `validation` accumulates on the host (trainer.py:411,547-554).
The metrics kernel writes straight into the send slot (e3_pack_inversion_record), the
collective is issued on the same stream, and nothing synchronises with the host.
"""
import torch
import torch.distributed as dist

from . import _lib

N_METRICS = 2  # mse, mae of the generated image against the target


def shard_range(n_images, rank, world_size):
    """Contiguous [lo, hi) of images owned by `rank` (the first n % world ranks get one more)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_images, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def record_length(n_latent):
    return 9 * 256 + n_latent * 512 + N_METRICS


def pack_records(w_plus, w_dec, image=None, target=None, out=None):
    """[B, record_length] fp32 rows, metrics computed on the device (CUDA only)."""
    lib = _lib.load()
    w_plus, w_dec = _lib.as_f32c(w_plus), _lib.as_f32c(w_dec)
    b, n_latent = w_plus.shape[0], w_dec.shape[1]
    if tuple(w_plus.shape[1:]) != (9, 256) or w_dec.shape[2] != 512:
        raise RuntimeError("pack_records: w_plus must be [B,9,256] and w_dec [B,n_latent,512]")
    if out is None:
        out = torch.empty(b, record_length(n_latent), device=w_plus.device, dtype=torch.float32)
    numel = 0
    if image is not None:
        image, target = _lib.as_f32c(image), _lib.as_f32c(target)
        numel = image[0].numel()
    _lib.check(lib.e3_pack_inversion_record(_lib.ptr(w_plus), _lib.ptr(w_dec), n_latent,
                                            _lib.ptr(image), _lib.ptr(target), b, numel,
                                            _lib.ptr(out), _lib.cur_stream()),
               "e3_pack_inversion_record")
    return out


def gather_records(local_records, group=None, out=None, equal_shards=False):
    """The single collective of the inversion pass: all-gather of the per-image records.

    equal_shards=True (weak scaling, same batch on every rank): exactly one
    all_gather_into_tensor — one NCCL kernel over NVLink, no host round trip.  Otherwise the
    per-rank counts are exchanged first and ragged shards are padded.  Works on any backend
    (gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_records
    world = dist.get_world_size(group)
    if world == 1:
        return local_records
    n_local, width = local_records.shape
    if equal_shards:
        counts = [n_local] * world
    else:
        mine = torch.tensor([n_local], device=local_records.device, dtype=torch.int64)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=group)
        counts = [int(c.item()) for c in every]
    if len(set(counts)) == 1:
        if out is None:
            out = local_records.new_empty(world * n_local, width)
        dist.all_gather_into_tensor(out, local_records.contiguous(), group=group)
        return out
    mx = max(counts)
    pad = local_records.new_zeros(mx, width)
    pad[:n_local] = local_records
    bufs = [local_records.new_empty(mx, width) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def unpack_records(records, n_latent):
    """-> (w_plus [N,9,256], w_dec [N,n_latent,512], metrics [N,K])"""
    a, b = 9 * 256, 9 * 256 + n_latent * 512
    return (records[:, :a].reshape(-1, 9, 256), records[:, a:b].reshape(-1, n_latent, 512),
            records[:, b:])
