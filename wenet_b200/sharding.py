"""Stream -> GPU partitioning for the multi-GPU path.

Streams are independent (no state is shared between two `struct FSK` instances, reference src/fsk.h:43-90,
or two deframers, src/drs232_ldpc.c:106-118), so the path shards with NO data-path collective: rank r owns a
contiguous block of streams, runs its own engine on its own GPU, and only the timing / packet counts are
reduced at the end.
"""


def shard_range(n_streams, rank, world):
    """Contiguous block [lo, hi) of `n_streams` global streams owned by `rank` of `world` (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(stream, n_streams, world):
    """Rank that owns global stream `stream` under shard_range()."""
    base, extra = divmod(n_streams, world)
    cut = extra * (base + 1)
    if stream < cut:
        return stream // (base + 1)
    return extra + (stream - cut) // base


def weak_global_streams(streams_per_gpu, world):
    """Weak scaling: per-GPU work fixed, the job grows with the number of GPUs."""
    return streams_per_gpu * world


def weighted_ranges(n_streams, weights):
    """Contiguous blocks [lo, hi) of `n_streams` global streams, one per engine, sized in proportion to `weights`
    (e.g. the host->device copy rate each GPU reaches when all of them copy at once: on an HGX box GPUs behind a busier
    host bridge get fewer streams, so that every engine finishes its feed at the same time).  Largest-remainder
    rounding: sizes add up to n_streams exactly and differ from the exact share by less than one stream."""
    w = [float(x) for x in weights]
    if not w or any(x < 0 for x in w) or sum(w) <= 0:
        raise ValueError("bad weights")
    tot = sum(w)
    exact = [n_streams * x / tot for x in w]
    size = [int(e) for e in exact]
    order = sorted(range(len(w)), key=lambda i: (exact[i] - size[i], -i), reverse=True)
    for i in order[:n_streams - sum(size)]:
        size[i] += 1
    out, lo = [], 0
    for sz in size:
        out.append((lo, lo + sz))
        lo += sz
    return out


def equal_ranges(n_streams, world):
    return [shard_range(n_streams, r, world) for r in range(world)]
