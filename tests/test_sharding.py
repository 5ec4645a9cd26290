"""N > 1 host logic on CPU: two processes over the gloo backend shard the characters of a small forest, score their
blocks (with the CPU oracle standing in for the device: the sharding logic is what is under test), all-reduce the
per-tree sums and must reproduce the unsharded totals."""
import os
import socket

import numpy as np
import pytest

from tests.helpers import random_string, tcm_matrix


def test_character_blocks_partition():
    from pcg_b200.sharding import character_block
    for n in (1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            blocks = [character_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    from pcg_b200.postorder import random_binary_trees
    rng = np.random.default_rng(5)
    n_trees, n_taxa, n_chars = 3, 9, 5
    children = random_binary_trees(n_trees, n_taxa, seed=2)
    leaves = np.stack([np.stack([random_string(rng, 40, amb=0.0) for _ in range(n_chars)]) for _ in range(n_taxa)])
    return children, leaves          # leaves: (n_taxa, n_chars, 40) uint8


def _score_block(children, leaves, lo, hi):
    """Per-tree sums over characters [lo, hi) by the CPU oracle's post-order."""
    from oracle import pyoracle as po
    n_trees, n_taxa = children.shape[0], leaves.shape[0]
    sub = np.ascontiguousarray(leaves[:, lo:hi, :])
    nc, L = sub.shape[1], sub.shape[2]
    out = np.zeros(n_trees, np.int64)
    if nc == 0:
        return out
    flat = np.ascontiguousarray(sub.reshape(-1))
    off = np.arange(n_taxa * nc, dtype=np.uint64) * np.uint64(L)
    ln = np.full(n_taxa * nc, L, np.uint32)
    _, col_cost, _ = po.cpu_bench_postorder(tcm_matrix("discrete"), children, n_taxa, nc, flat, off, ln, 0,
                                            n_trees * nc, 1, "port")
    return col_cost.reshape(n_trees, nc).sum(axis=1)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from pcg_b200.sharding import allreduce_tree_costs, character_block
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    children, leaves = _problem()
    lo, hi = character_block(leaves.shape[1], rank, world)
    part = torch.from_numpy(_score_block(children, leaves, lo, hi))
    allreduce_tree_costs(part, dist)
    if rank == 0:
        q.put(part.numpy().tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_character_sharding_gloo(oracle_mod):
    import torch.multiprocessing as mp
    children, leaves = _problem()
    want = _score_block(children, leaves, 0, leaves.shape[1]).tolist()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want
