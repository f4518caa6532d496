"""CPU: the host-side logic of the multi-GPU path -- joining the ranks' flat outputs -- including the
padded all-gather bench.py uses, on the gloo backend with world_size 2."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest

from spectralclustersupertree_b200.engine import merge_sharded
from spectralclustersupertree_b200.scs import _tree_from_flat

NAMES = list("abcdefgh")


def example_parts():
    """Root with two sub-problem slots (nodes 1 and 2); rank 0 solves slot 1, rank 1 solves slot 2."""
    prefix = 3
    # rank 0: slot 1 -> ((a,b),c)
    p0 = np.array([-1, 0, 0, 1, 3, 3, 1], dtype=np.int32)
    t0 = np.array([-1, -1, -1, -1, 0, 1, 2], dtype=np.int32)
    # rank 1: slot 2 -> (d,(e,f))
    p1 = np.array([-1, 0, 0, 2, 2, 4, 4], dtype=np.int32)
    t1 = np.array([-1, -1, -1, 3, -1, 4, 5], dtype=np.int32)
    return [(p0, t0, prefix), (p1, t1, prefix)]


def test_merge_sharded_concatenates_past_the_prefix():
    parent, taxon = merge_sharded(example_parts())
    assert len(parent) == 3 + 4 + 4
    assert (parent[1:] < np.arange(1, len(parent))).all()
    tree = _tree_from_flat(parent, taxon, NAMES)
    assert tree.clade_sets() >= {frozenset("ab"), frozenset("abc"), frozenset("ef"), frozenset("def")}
    assert sorted(tree.get_tip_names()) == list("abcdef")


def test_merge_sharded_keeps_a_tip_placed_in_a_shared_slot():
    prefix = 3
    p0 = np.array([-1, 0, 0, 1, 1], dtype=np.int32)
    t0 = np.array([-1, -1, -1, 0, 1], dtype=np.int32)  # slot 1 -> (a,b); slot 2 belongs to rank 1
    p1 = np.array([-1, 0, 0], dtype=np.int32)
    t1 = np.array([-1, -1, 2], dtype=np.int32)  # slot 2 resolved to the single tip c
    parent, taxon = merge_sharded([(p0, t0, prefix), (p1, t1, prefix)])
    tree = _tree_from_flat(parent, taxon, NAMES)
    assert sorted(tree.get_tip_names()) == list("abc")


def test_merge_sharded_rejects_different_prefixes():
    parts = example_parts()
    parts[1] = (parts[1][0], parts[1][1], 4)
    with pytest.raises(ValueError):
        merge_sharded(parts)


def _worker(rank: int, world: int, port: int, queue) -> None:
    import torch
    import torch.distributed as dist

    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parent, taxon, prefix = example_parts()[rank]
        size = torch.tensor([len(parent), prefix], dtype=torch.int64)
        sizes = [torch.zeros_like(size) for _ in range(world)]
        dist.all_gather(sizes, size)
        longest = int(max(int(s[0]) for s in sizes))
        mine = torch.full((2, longest), -2, dtype=torch.int32)
        mine[0, : len(parent)] = torch.from_numpy(parent)
        mine[1, : len(taxon)] = torch.from_numpy(taxon)
        everyone = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)
        parts = [(everyone[r][0, : int(sizes[r][0])].numpy(), everyone[r][1, : int(sizes[r][0])].numpy(),
                  int(sizes[r][1])) for r in range(world)]  # fmt: skip
        merged_parent, merged_taxon = merge_sharded(parts)
        queue.put((rank, merged_parent.tolist(), merged_taxon.tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_all_gather_and_merge_world_size_2():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expected = merge_sharded(example_parts())
    for _, parent, taxon in results:
        assert parent == expected[0].tolist()
        assert taxon == expected[1].tolist()
