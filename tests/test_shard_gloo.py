"""CPU, world_size 2 over gloo: the multi-GPU plumbing of the row-striped index (go-vectorsearch_b200/shard.py).

Each rank owns the rows `rank, rank+world, ...` of every posting list, answers the query on its shard,
all-gathers the shard-local top-k and merges.  On the GPU box the local search and the merge are libvscuda
kernels; here the oracle stands in for the local search so that the host logic (striping covers every row
exactly once, gather layout, merge order and dedup rule) is checked without a GPU.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_np(g_ids, g_sims, g_counts, k):
    """Reference merge: similarity desc (float32), id asc, one hit per id."""
    out = []
    for q in range(g_ids.shape[1]):
        cand = []
        for g in range(g_ids.shape[0]):
            for j in range(int(g_counts[g, q])):
                cand.append((-float(g_sims[g, q, j]), int(g_ids[g, q, j])))
        cand.sort()
        seen, hits = set(), []
        for negs, i in cand:
            if i not in seen:
                seen.add(i)
                hits.append((i, np.float32(-negs)))
            if len(hits) == k:
                break
        out.append(hits)
    return out


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        import oracle
        from conftest import load_pkg
        from _util import unit_rows
        pkg = load_pkg()
        dist.init_process_group("gloo", rank=rank, world_size=world)
        n, d, C, nprobe, k, nq = 3000, 64, 12, 4, 10, 5
        rows = oracle.quantize_matrix_f32(unit_rows(n, d, 1))
        cent = oracle.quantize_matrix_f32(unit_rows(C, d, 2))
        _, lists = oracle.argmax_MxN(cent, rows)
        lists = lists.astype(np.uint32)
        doc = (np.arange(n) // 2).astype(np.uint64)           # two embeddings per document, split across ranks
        qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 3))
        mine = pkg.shard.stripe(n, rank, world)
        assert len(mine) == pkg.shard.local_count(n, rank, world)
        ids = np.zeros((nq, k), np.uint64)
        sims = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.int32)
        for i in range(nq):
            li, ls = oracle.search(qs[i], cent, rows[mine], lists[mine], doc[mine], nprobe, k)
            ids[i, :len(li)], sims[i, :len(li)], counts[i] = li, ls, len(li)
        g = pkg.shard.gather_hits(torch.from_numpy(ids.view(np.int64)), torch.from_numpy(sims), torch.from_numpy(counts))
        g_ids, g_sims, g_counts = g[0].numpy().view(np.uint64), g[1].numpy(), g[2].numpy()
        assert g_ids.shape == (world, nq, k)
        # every rank's stripe is disjoint and together they cover all rows
        cover = torch.zeros(n, dtype=torch.int32)
        cover[torch.from_numpy(mine)] = 1
        dist.all_reduce(cover)
        assert int(cover.min()) == 1 and int(cover.max()) == 1
        merged = _merge_np(g_ids, g_sims, g_counts, k)
        for i in range(nq):
            wi, ws = oracle.search(qs[i], cent, rows, lists, doc, nprobe, k)
            assert [h[0] for h in merged[i]] == wi.tolist(), f"rank {rank} query {i}"
            assert np.array([h[1] for h in merged[i]], np.float32).view(np.uint32).tolist() == ws.view(np.uint32).tolist()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error: " + repr(e) + "\n" + traceback.format_exc()))


def test_striped_search_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(r[1] == "ok" for r in res), res


def test_stripe_partition():
    from conftest import load_pkg
    pkg = load_pkg()
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            parts = [pkg.shard.stripe(n, r, world) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert [len(p) for p in parts] == [pkg.shard.local_count(n, r, world) for r in range(world)]


def test_upload_keeps_the_stripe_rule():
    """After uploads every rank still owns exactly the primary keys with key % world == rank."""
    from conftest import load_pkg
    pkg = load_pkg()
    for world in (1, 2, 3, 8):
        n = 1003
        owned = [set(pkg.shard.stripe(n, r, world).tolist()) for r in range(world)]
        for n_new in (1, 5, 64, 0):
            parts = [pkg.shard.upload_stripe(n, n_new, r, world) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n_new))
            for r in range(world):
                owned[r].update(int(n + i) for i in parts[r])
            n += n_new
        for r in range(world):
            assert owned[r] == set(pkg.shard.stripe(n, r, world).tolist())


# ---- k-means over contiguous row blocks: the relay of shard.kmeans_step_relay ------------------------------------------
def _kmeans_worker(rank, world, port, q):
    """On the GPU box assign / accumulate / finish are libvscuda calls (vs_argmax_MxN_dev, vs_kmeans_accumulate_dev,
    vs_kmeans_finish_dev); here numpy stand-ins restate them (float32 sums continued in row order, k_means.go:80-99) so
    that the relay's control flow -- who waits for whom, what is broadcast -- is checked on CPU against the oracle's
    single-process iteration."""
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        import oracle
        from conftest import load_pkg
        from _util import unit_rows
        pkg = load_pkg()
        dist.init_process_group("gloo", rank=rank, world_size=world)
        n, d, k = 1501, 32, 6
        data = oracle.quantize_matrix_f32(unit_rows(n, d, 11))
        cent = data[np.random.default_rng(3).choice(n, k, replace=False)].copy()
        cent[4] = cent[2]                                   # an empty cluster keeps its previous mean
        lo, hi = pkg.shard.block_range(n, rank, world)
        mine = data[lo:hi]
        means = np.full((k, d), 0.5, np.float32)            # state of the finishing rank
        means_o = means.copy()
        cent_o = cent.copy()
        cent_t = torch.from_numpy(cent.copy())
        sums = torch.zeros(k * d, dtype=torch.float32)
        counts = torch.zeros(k, dtype=torch.int64)
        state = {}
        for it in range(3):
            a_o, c_o, new_o, conv_o = oracle.kmeans_step(data, cent_o, means_o)
            cur = cent_t.numpy().copy()

            def assign():
                state["a"] = oracle.argmax_MxN(cur, mine)[1]

            def accumulate(s, c):
                sv = s.numpy().reshape(k, d)
                x = oracle.dequantize_matrix_f32(mine)
                for i, j in enumerate(state["a"]):           # row order; float32 additions (k_means.go:81-86)
                    sv[j] += x[i]
                    c[j] += 1

            def finish(s, c):
                sv = s.numpy().reshape(k, d)
                for j in range(k):
                    if int(c[j]) > 0:
                        means[j] = sv[j] / np.float32(int(c[j]))
                newc = oracle.quantize_matrix_f32(means)
                conv = bool((newc[:, 8:] == cur[:, 8:]).all())
                return torch.from_numpy(newc), conv

            conv = pkg.shard.kmeans_step_relay(assign, accumulate, finish, sums, counts, cent_t)
            assert (state["a"] == a_o[lo:hi]).all()
            assert (cent_t.numpy() == new_o).all(), f"rank {rank} iteration {it}"
            assert conv == conv_o
            if rank == world - 1:
                assert (counts.numpy() == c_o).all()
                assert (means.view(np.uint32) == means_o.view(np.uint32)).all()
            cent_o = new_o
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error: " + repr(e) + "\n" + traceback.format_exc()))


def test_kmeans_relay_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_kmeans_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(r[1] == "ok" for r in res), res


def test_block_partition():
    from conftest import load_pkg
    pkg = load_pkg()
    for n in (0, 1, 7, 1000, 1501):
        for world in (1, 2, 3, 8):
            r = [pkg.shard.block_range(n, i, world) for i in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


# ---- the probe stage shared between the ranks: shard.SharedProbes -------------------------------------------------------
def _probe_worker(rank, world, port, q):
    """On the GPU box Index.ProbeDev is vs_probe_dev; here the oracle's select_probes writes into the same buffers, so the
    gather layout (rank-major = query order), the status words and the share each rank takes are checked without a GPU."""
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import ctypes as C
        import torch
        import torch.distributed as dist
        import oracle
        from conftest import load_pkg
        from _util import unit_rows
        pkg = load_pkg()
        dist.init_process_group("gloo", rank=rank, world_size=world)
        d, Cn, nprobe, nq = 64, 20, 5, 12
        cent = oracle.quantize_matrix_f32(unit_rows(Cn, d, 2))
        qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 3))

        class FakeIndex:
            def ProbeDev(self, q_slice, npb, d_probe, d_status, ctx=None):
                out = np.ctypeslib.as_array(C.cast(d_probe, C.POINTER(C.c_int32)), shape=(len(q_slice) * npb,))
                st = np.ctypeslib.as_array(C.cast(d_status, C.POINTER(C.c_int32)), shape=(len(q_slice),))
                for i, row in enumerate(q_slice):
                    p, _ = oracle.select_probes(row, cent, npb)
                    out[i * npb:(i + 1) * npb] = p
                    st[i] = 1 if (rank == 1 and i == 0) else 0          # one flagged query, owned by rank 1

        sp = pkg.shard.SharedProbes(nq, nprobe, torch.device("cpu"), world, rank)
        assert sp.rows_of() == slice(rank * nq // world, (rank + 1) * nq // world)
        status = torch.full((nq,), 7, dtype=torch.int32)
        probe = sp.select_and_gather(FakeIndex(), qs[sp.rows_of()], status, ctx=None).numpy().reshape(nq, nprobe)
        for i in range(nq):
            want, _ = oracle.select_probes(qs[i], cent, nprobe)
            assert probe[i].tolist() == want.tolist(), f"rank {rank} query {i}"
        want_status = [0] * nq
        want_status[nq // world] = 1
        assert status.tolist() == want_status
        with pytest.raises(ValueError):
            pkg.shard.SharedProbes(nq + 1, nprobe, torch.device("cpu"), world, rank)
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error: " + repr(e) + "\n" + traceback.format_exc()))


def test_shared_probe_stage_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_probe_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(r[1] == "ok" for r in res), res
