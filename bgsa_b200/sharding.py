"""Contiguous subject ranges per device -- the host-side mirror of dispatch_task()
(original/BGSA_AVX512/global.c:86-, original/BGSA_KNC/global.c:374-): every device gets an equal
share rounded down to the padding unit, the last device takes the remainder.  The reference's unit
is the SIMD width (*_V_NUM); ours is the 32-subject tile.  No collective is involved on the data
path: scores are gathered to the host, device-major, exactly as cal_mic.c:470-471 lays them out.
"""
from __future__ import annotations

UNIT = 32


def shard_counts(total: int, world: int, unit: int = UNIT) -> list[int]:
    if world < 1 or total < 0:
        raise ValueError("bad shard request")
    per = (total // world) // unit * unit
    counts = [per] * world
    counts[-1] = total - per * (world - 1)
    return counts


def shard_range(total: int, rank: int, world: int, unit: int = UNIT) -> tuple[int, int]:
    """(first, count) of `rank`'s contiguous subject range."""
    counts = shard_counts(total, world, unit)
    return sum(counts[:rank]), counts[rank]
