"""Sharding of call legs over the GPUs of one box: legs are independent,
so leg s simply lives on rank s % world; each rank owns one engine, no collective on the
data path.  Used by bench.py (one process per GPU) and by the world_size-2 gloo test."""


def legs_of_rank(total_legs, rank, world):
    """Global leg ids served by `rank`."""
    return list(range(rank, total_legs, world))


def rank_of_leg(leg, world):
    return leg % world


def merge_outputs(per_rank_outputs, world):
    """Inverse of legs_of_rank for gathered per-rank [legs_of_rank][...] arrays: returns the list in
    global leg order."""
    total = sum(len(o) for o in per_rank_outputs)
    out = [None] * total
    for r, outs in enumerate(per_rank_outputs):
        for j, o in enumerate(outs):
            out[r + j * world] = o
    return out
