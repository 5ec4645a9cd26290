"""How the DO path is split over N processes (one per GPU) — host logic only.

The path shards without any data-path exchange: pair batches are split evenly; tree scoring is split by
CHARACTER BLOCK — every rank keeps all node strings of its characters resident and runs every level locally, because
a (tree, character) column never interacts with another (per character independent evaluation,
Bio/Sequence/Block/Character.hs:440-463 hexZipMeta; tree cost = sum over characters, Bio/Sequence/Block.hs:87-106).
The only collective is one all-reduce of the int64[n_trees] partial sums.
"""
from __future__ import annotations

from typing import Tuple


def character_block(n_chars: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the characters rank `rank` owns: contiguous blocks, sizes differing by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    q, r = divmod(n_chars, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def pair_block(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of a pair batch for rank `rank` (strong scaling of a fixed batch)."""
    return character_block(n_pairs, rank, world)


def allreduce_tree_costs(costs, dist=None):
    """Sum the per-tree partial costs of all ranks in place (torch tensor, int64[n_trees]); the path's only
    collective.  No-op for a single process."""
    if dist is None:
        import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(costs, op=dist.ReduceOp.SUM)
    return costs
