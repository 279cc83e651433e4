"""CPU: bench.py's synthetic store is the same for every world size -- a rank's rows are the stripe `row mod world ==
rank` of the one-GPU store -- although a rank only generates the lanes it owns."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("count", [64, 61, 9, 8, 3, 1])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_rank_rows_are_a_stripe_of_the_store(count, world):
    dev = torch.device("cpu")
    for ci, r0 in ((0, 0), (5, 5 * bench.CHUNK)):
        full, first = bench.store_chunk_rows(torch, ci, r0, count, 0, 1, dev)
        assert first == 0 and full.shape == (count, bench.D)
        assert torch.allclose(full.norm(dim=1), torch.ones(count), atol=1e-5)
        owned = 0
        for rank in range(world):
            part, first = bench.store_chunk_rows(torch, ci, r0, count, rank, world, dev)
            assert first == (rank - r0) % world
            if part is None:
                assert first >= count
                continue
            assert torch.equal(part, full[first::world])
            owned += part.shape[0]
        assert owned == count


def test_chunks_differ():
    dev = torch.device("cpu")
    a, _ = bench.store_chunk_rows(torch, 0, 0, 16, 0, 1, dev)
    b, _ = bench.store_chunk_rows(torch, 1, bench.CHUNK, 16, 0, 1, dev)
    assert not torch.equal(a, b) and bench.CHUNK % bench.STORE_LANES == 0
