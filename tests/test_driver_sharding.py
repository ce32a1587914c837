"""The N>1 path on CPU: world_size-2 gloo processes shard pairs, the oracle stands in for the GPU hot path (it is
the checker here, the thing under test is the partition + the single gather)."""
import os

import pytest

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pope_b200 import driver


def test_shard_range_partitions_exactly():
    for n in (1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            spans = [driver.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flatten_slots_orders_by_pair_then_row():
    out = {"i_ids": torch.tensor([[3, 9, 0], [1, 0, 0]]), "j_ids": torch.tensor([[5, 2, 0], [7, 0, 0]]),
           "mconf": torch.tensor([[.5, .6, 0], [.7, 0, 0]]), "mkpts0_f": torch.zeros(2, 3, 2), "mkpts1_f": torch.ones(2, 3, 2),
           "counts": torch.tensor([2, 1], dtype=torch.int32)}
    flat = driver.flatten_slots(out, pair_offset=10)
    assert flat["b_ids"].tolist() == [10, 10, 11] and flat["i_ids"].tolist() == [3, 9, 1] and flat["j_ids"].tolist() == [5, 2, 7]


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pope_oracle as O
    from pope_b200 import synth
    hw = (8, 10)

    def local_fn(lo, hi):
        f0, f1 = synth.coarse_features(5, n_pairs, 80, 80, 64, sigma=0.8)      # every rank regenerates, slices its shard
        ff0, ff1 = synth.fine_feature_maps(6, n_pairs, 32, 40, 128, channels_last=False)
        if hi == lo:
            e = torch.empty(0)
            return {"b_ids": e.long(), "i_ids": e.long(), "j_ids": e.long(), "mconf": e, "mkpts0_f": torch.empty(0, 2),
                    "mkpts1_f": torch.empty(0, 2)}
        o = O.match_pairs(f0[lo:hi], f1[lo:hi], ff0[lo:hi], ff1[lo:hi], (64, 80), hw, hw)
        o["b_ids"] = o["b_ids"] + lo
        return o

    got = driver.run_sharded(n_pairs, rank, world, local_fn)
    # the job-level gather used by bench.py for N > 1: two "steps" per rank, one collective at the end
    lo, hi = driver.shard_range(n_pairs, rank, world)
    loc = local_fn(lo, hi)
    m = loc["i_ids"].numel()
    cap = 2 * m + 8
    res = {"b_ids": torch.zeros(cap, dtype=torch.int64), "i_ids": torch.zeros(cap, dtype=torch.int64),
           "j_ids": torch.zeros(cap, dtype=torch.int64), "mconf": torch.zeros(cap), "mkpts0_f": torch.zeros(cap, 2),
           "mkpts1_f": torch.zeros(cap, 2), "n_pairs": hi - lo, "counts": torch.zeros(hi - lo + 2, dtype=torch.int32)}
    for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_f", "mkpts1_f"):
        res[k][:m] = loc[k] if k != "b_ids" else loc[k] - lo
    res["counts"][hi - lo] = m
    job = driver.JobGather(2, cap, torch.device("cpu"), rank=rank, world=world)
    job.add(res, lo)
    job.add(res, lo)
    recs, sizes = job.finish()
    # the compact 20-byte records (mkpts0_f implied by i) through the same gather, with the per-rank checksums bench.py uses
    jobc = driver.JobGather(2, cap, torch.device("cpu"), rank=rank, world=world, compact=True)
    jobc.add(res, lo)
    jobc.add(res, lo)
    mine = jobc.checksum()
    sums = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sums, mine)
    recs_c, sizes_c = jobc.finish()
    # retrieval sharded over the ranks: each rank scores its block of the crops, one all-gather of the scores, then the running
    # top-k over all of them (the oracle stands in for the two CUDA kernels)
    from pope_b200 import retrieval
    q_tok, ref_tok = synth.retrieval_tokens(9, 37, 48)
    rlo, rhi = driver.shard_range(37, rank, world)
    ret = retrieval.retrieve_topk_sharded(O.cosine_scores(q_tok, ref_tok[rlo:rhi]), 37, rank, world, 3,
                                          topk_fn=lambda sc, k: O.running_topk(sc.tolist(), k))
    if rank == 0:
        # numpy arrays travel through the queue by value; tensors would travel as file descriptors that die with this process
        payload = {k: (v.numpy().copy() if torch.is_tensor(v) else v) for k, v in got.items()}
        payload["job_sizes"] = sizes
        payload["retrieval"] = (ret[0].numpy().copy(), ret[1], ret[2])
        payload["job_b"] = torch.cat([recs[r, :sizes[r], 0] for r in range(world)]).numpy().copy()
        payload["job_conf"] = torch.cat([recs[r, :sizes[r], 3] for r in range(world)]).view(torch.float32).numpy().copy()
        assert sizes_c == sizes and recs_c.shape[2] == 5
        assert torch.equal(jobc.checksum(recs_c, sizes_c), torch.cat(sums))
        for r in range(world):
            full = driver.unpack_records(recs[r, :sizes[r]])
            comp = driver.unpack_records(recs_c[r, :sizes[r]], hw[1], 8.0)
            for k in full:
                assert torch.equal(full[k], comp[k]), k
        q.put(payload)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    from oracle import pope_oracle as O
    from pope_b200 import synth
    n_pairs, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:          # a free port chosen by the kernel (fixed ports collide between test runs)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    got = {k: (torch.from_numpy(v) if hasattr(v, "dtype") else v) for k, v in got.items()}      # (tuples pass through)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    f0, f1 = synth.coarse_features(5, n_pairs, 80, 80, 64, sigma=0.8)
    ff0, ff1 = synth.fine_feature_maps(6, n_pairs, 32, 40, 128, channels_last=False)
    want = O.match_pairs(f0, f1, ff0, ff1, (64, 80), (8, 10), (8, 10))
    assert want["b_ids"].numel() > 20
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(got[k], want[k]), k
    assert torch.equal(got["mconf"], want["mconf"])
    assert torch.equal(got["mkpts1_f"], want["mkpts1_f"]) and torch.equal(got["mkpts0_f"], want["mkpts0_f"])
    assert sum(got["per_rank_matches"]) == want["b_ids"].numel()
    # sharded retrieval: the same scores, slot scores and slot indices as the one-process loop over all 37 crops
    q_tok, ref_tok = synth.retrieval_tokens(9, 37, 48)
    all_scores = O.cosine_scores(q_tok, ref_tok)
    want_s, want_i = O.running_topk(all_scores.tolist(), 3)
    r_scores, r_slot_s, r_slot_i = got["retrieval"]
    assert torch.equal(torch.as_tensor(r_scores), all_scores) and r_slot_i == want_i and r_slot_s == want_s and min(want_i) >= 0
    # job-level gather: each rank contributed its matches twice (two identical steps)
    assert got["job_sizes"] == [2 * x for x in got["per_rank_matches"]]
    per = torch.tensor(got["per_rank_matches"])
    off = torch.cat([torch.zeros(1, dtype=torch.long), per.cumsum(0)])
    want_b = torch.cat([want["b_ids"][off[r]:off[r + 1]].repeat(2) for r in range(2)]).to(torch.int32)
    want_c = torch.cat([want["mconf"][off[r]:off[r + 1]].repeat(2) for r in range(2)])
    assert torch.equal(got["job_b"], want_b) and torch.equal(got["job_conf"], want_c)


def test_bgr_to_gray_equals_cv2_for_every_colour():
    """The grey conversion of the evaluation loop (eval_linemod_json.py:103, :109) on the device = cv2's 8-bit path."""
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    from pope_b200 import driver
    r = np.arange(256, dtype=np.uint8)
    for b in range(0, 256, 8):                               # 32 slabs of 8 x 256 x 256 colours = all 2^24
        bb = np.arange(b, b + 8, dtype=np.uint8)
        cube = np.stack(np.meshgrid(bb, r, r, indexing="ij"), -1).reshape(8, 65536, 3)
        assert np.array_equal(driver.bgr_to_gray(torch.from_numpy(cube)).numpy(), cv2.cvtColor(cube, cv2.COLOR_BGR2GRAY)), b
