"""Batch sharding of utterances over ranks (one process per GPU, no data-path
collective - the reference's analogue is one gunicorn worker per GPU,
gunicorn_config.py:43-60).  torch.distributed is used only for the benchmark's
barrier and max-over-ranks timing."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """contiguous, balanced [lo, hi) slice of `n_items` utterances for `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def split_chunks(n_frames, chunk, halo=34):
    """Long-audio chunking along time: [(start, stop, keep_lo, keep_hi)] in mel frames.
    The generator's receptive field is +-34 mel frames, so a chunk
    vocoded with `halo` extra frames each side is exact on its kept part."""
    out = []
    pos = 0
    while pos < n_frames:
        end = min(n_frames, pos + chunk)
        s, e = max(0, pos - halo), min(n_frames, end + halo)
        out.append((s, e, pos - s, end - s))
        pos = end
    return out


def max_over_ranks(x, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def vocode_sharded(model, mel, rank, world):
    """this rank's slice of a global batch of mels -> (lo, hi, wav slice)."""
    lo, hi = shard_range(mel.shape[0], rank, world)
    return lo, hi, model(mel[lo:hi].contiguous())
