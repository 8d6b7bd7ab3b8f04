"""Multi-GPU trace commitment: the shard plan, the column-sharded -> all-to-all -> row-sharded flow under a
real world_size-2/4 gloo process group on CPU (kernels replaced by the oracle double), and -- on the B200 -- the CUDA
stage kernels driven through the same plan with the ranks emulated on one GPU."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
import starky_bls12_381_b200 as sb
from helpers import P, random_trace
from starky_bls12_381_b200.sharded import commit_sharded, shard_plan


def test_shard_plan_ragged_and_limits():
    pl = shard_plan(97330, 10, 1, 8)           # MillerLoop on 8 GPUs
    assert sum(pl.col_count) == 97330 and pl.col_count == (12167, 12167) + (12166,) * 6
    assert pl.col_start[0] == 0 and all(pl.col_start[g + 1] == pl.col_start[g] + pl.col_count[g] for g in range(7))
    assert pl.rows_per_rank == 256 and pl.n_lde == 2048
    assert pl.send_splits(0) == [12167 * 256] * 8 and pl.recv_splits(3) == [c * 256 for c in pl.col_count]
    # what one rank sends is what the others expect from it
    for g in range(8):
        for h in range(8):
            assert pl.send_splits(g)[h] == pl.recv_splits(h)[g]
    assert pl.a2a_bytes_out(2) == 8 * 12166 * 256 * 7
    fe = shard_plan(73527, 13, 2, 4)           # FinalExp on 4 GPUs: LDE bytes x 3/4 cross the fabric
    assert sum(fe.a2a_bytes_out(g) for g in range(4)) == 8 * 73527 * 32768 * 3 // 4
    assert shard_plan(5, 4, 1, 1).col_count == (5,)
    with pytest.raises(ValueError):
        shard_plan(100, 4, 1, 3)               # row blocks are power-of-two slices
    with pytest.raises(ValueError):
        shard_plan(100, 4, 1, 2 ** 4)          # < 32 positions per rank


def _worker(rank, world, init_file, n_cols, log_n, rate_bits, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sharded_double import OracleBackend
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        trace = random_trace(np.random.default_rng(4242), n_cols, log_n, full_width=True)     # same seed on every rank
        plan = shard_plan(n_cols, log_n, rate_bits, world)
        c0, cg = plan.col_start[rank], plan.col_count[rank]
        out = commit_sharded(OracleBackend(), plan, rank, trace[c0:c0 + cg])
        q.put((rank, out["cap"], out["digests"].numpy().view(np.uint64).copy(), out["rows"].numpy().view(np.uint64).copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_cols,log_n,rate_bits", [(2, 13, 5, 1), (4, 9, 5, 2), (2, 3, 6, 1)])
def test_sharded_commit_over_gloo_matches_single_process_oracle(world, n_cols, log_n, rate_bits):
    from helpers import pos_to_leaf
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rendezvous")
        procs = [ctx.Process(target=_worker, args=(r, world, init_file, n_cols, log_n, rate_bits, q)) for r in range(world)]
        for pr in procs:
            pr.start()
        results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
        for pr in procs:
            pr.join(timeout=60)
            assert pr.exitcode == 0
    trace = random_trace(np.random.default_rng(4242), n_cols, log_n, full_width=True)
    want = O.lde_commit(O.make_params(log_n=log_n, n_cols=n_cols, rate_bits=rate_bits), trace)
    perm = pos_to_leaf(log_n, rate_bits)
    lde_pos = want["leaves"][perm].T                                   # [C][N] in device position order
    rb = (1 << (log_n + rate_bits)) // world
    for rank, cap, digests, rows in results:
        assert np.array_equal(cap, want["cap"])                        # every rank ends with the reference's cap
        assert np.array_equal(digests, want["digests"][perm])          # gathered digests, position order
        assert np.array_equal(rows, lde_pos[:, rank * rb:(rank + 1) * rb])   # its row block holds ALL columns, in order


# ---------------------------------------------------------------- CUDA stage kernels, ranks emulated on one GPU
@pytest.mark.gpu
@pytest.mark.parametrize("world,n_cols,log_n,rate_bits", [(1, 20, 6, 1), (2, 37, 6, 1), (4, 45, 7, 2), (8, 203, 10, 2)])
def test_gpu_sharded_stage_kernels_match_unsharded_commit(world, n_cols, log_n, rate_bits):
    from starky_bls12_381_b200.sharded import GpuBackend
    ctx = sb.Context(0)
    try:
        p = sb.Params(sb.StarkId.CUSTOM, log_n, n_cols, 0, 3, rate_bits, 4, 2, 16, 84, 4, 5, 0, 0, 0)
        trace = random_trace(np.random.default_rng(77), n_cols, log_n, full_width=True)
        base = ctx.lde_commit(p, trace)                                 # the single-GPU path (itself checked against the oracle)
        plan = shard_plan(n_cols, log_n, rate_bits, world)
        be = GpuBackend(ctx, p)
        slabs = [be.lde_cols(plan, g, trace[plan.col_start[g]:plan.col_start[g] + plan.col_count[g]]) for g in range(world)]
        rb = plan.rows_per_rank
        digs = []
        for h in range(world):                                          # what the all-to-all delivers to rank h
            parts = [slabs[g].view(world, plan.col_count[g], rb)[h].reshape(-1) for g in range(world)]
            rows = torch.cat(parts).view(n_cols, rb)
            assert np.array_equal(rows.cpu().numpy().view(np.uint64), base["lde"][:, h * rb:(h + 1) * rb])
            digs.append(be.hash_rows(plan, rows.contiguous()))
        cap = be.merkle_cap(plan, torch.cat(digs, dim=0))
        assert np.array_equal(cap, base["cap"])
    finally:
        ctx.close()


# ---------------------------------------------------------------- row-sharded quotient (SURVEY 8e phase 2)
def test_successor_block():
    from starky_bls12_381_b200.sharded import successor_block
    fe8 = shard_plan(73527, 13, 2, 8)          # FinalExp on 8 GPUs: 4 cosets of 8192, blocks of 4096 = half a coset
    assert [successor_block(fe8, r) for r in range(8)] == [1, 0, 3, 2, 5, 4, 7, 6]
    fe4 = shard_plan(73527, 13, 2, 4)          # one coset per rank: no halo
    assert [successor_block(fe4, r) for r in range(4)] == [None] * 4
    ml8 = shard_plan(97330, 10, 1, 8)          # MillerLoop: 2 cosets of 1024, blocks of 256
    assert [successor_block(ml8, r) for r in range(8)] == [1, 2, 3, 0, 5, 6, 7, 4]


def _q_worker(rank, world, init_file, log_n, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import toy_air
    from sharded_double import OracleBackend
    from starky_bls12_381_b200.sharded import quotient_sharded
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        with tempfile.TemporaryDirectory() as d:
            air = toy_air.limbs(d, 4)
            trace, pis = air["witness"](log_n)
            p = O.make_params(n_cols=air["n_cols"], n_pis=air["n_pis"], degree=air["degree"], rate_bits=air["rate_bits"], log_n=log_n)
            plan = shard_plan(air["n_cols"], log_n, air["rate_bits"], world)
            be = OracleBackend(air["flat"], p, trace)
            c0, cg = plan.col_start[rank], plan.col_count[rank]
            com = commit_sharded(be, plan, rank, trace[c0:c0 + cg])
            out = quotient_sharded(be, plan, rank, com["rows"], com["cap"], pis)
            q.put((rank, out["alphas"].copy(), out["q"].numpy().view(np.uint64).copy()))
            dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n", [(2, 6), (4, 6), (8, 7)])
def test_sharded_quotient_over_gloo_matches_single_process_oracle(world, log_n):
    """world 2: a block = two cosets (rate_bits 2 -> 4 cosets), no halo; world 4: one coset per rank; world 8: half a
    coset per rank -> the halo all-gather is exercised and the double asserts every rank received its successor's row."""
    import toy_air
    from helpers import pos_to_natural
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rendezvous")
        procs = [ctx.Process(target=_q_worker, args=(r, world, init_file, log_n, q)) for r in range(world)]
        for pr in procs:
            pr.start()
        results = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
        for pr in procs:
            pr.join(timeout=60)
            assert pr.exitcode == 0
        air = toy_air.limbs(d, 4)
        trace, pis = air["witness"](log_n)
        p = O.make_params(n_cols=air["n_cols"], n_pis=air["n_pis"], degree=air["degree"], rate_bits=air["rate_bits"], log_n=log_n)
        cap = O.lde_commit(p, trace)["cap"]
        alphas = O.challenger_run(cap.reshape(-1), 2)
        want = O.quotient_values(air["flat"], p, trace, pis, alphas)[:, pos_to_natural(log_n, air["rate_bits"])]
    for rank, al, qv in results:
        assert np.array_equal(al, alphas)
        assert np.array_equal(qv, want)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_gpu_row_sharded_quotient_matches_unsharded(world):
    """sb_quotient_rows_device on every row block (ranks emulated on one GPU, halo rows taken from the successor block)
    == sb_quotient_values of the single-GPU path, for a real constraint program (MillerLoop, degree 3, rate_bits 1)."""
    from helpers import pos_to_natural
    from starky_bls12_381_b200 import airfiles
    from starky_bls12_381_b200.sharded import GpuBackend, successor_block
    name, log_n = "miller_loop", 8
    info = sb.STARKS[name]
    airfiles.air_path(name, "airbin")
    ctx = sb.Context(0)
    try:
        p = sb.standard_params(info.stark_id, log_n)
        rng = np.random.default_rng(99)
        trace = random_trace(rng, info.columns, log_n)
        pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
        base = ctx.lde_commit(p, trace)
        be = GpuBackend(ctx, p)
        alphas = be.alphas(base["cap"])
        want = ctx.quotient_values(p, pis, alphas)[:, pos_to_natural(log_n, p.rate_bits)]     # position order
        plan = shard_plan(info.columns, log_n, p.rate_bits, world)
        rb = plan.rows_per_rank
        lde = torch.from_numpy(base["lde"].view(np.int64)).cuda()
        blocks = [lde[:, h * rb:(h + 1) * rb].contiguous() for h in range(world)]
        for h in range(world):
            nb = successor_block(plan, h)
            halo = blocks[nb][:, 0].contiguous() if nb is not None else None
            got = be.quotient_rows(plan, h, blocks[h], halo, pis, alphas)
            ctx.synchronize()
            assert np.array_equal(got.cpu().numpy().view(np.uint64), want[:, h * rb:(h + 1) * rb]), h
    finally:
        ctx.close()


# ---------------------------------------------------------------- the whole proof, sharded (SURVEY 8e small collectives)
def _prove_sharded_threads(world, p, trace, pis, airbin=None, stark_id=None, fused=False):
    """`world` host threads, one sb_ctx each on the one GPU, ThreadGroup collectives: the control flow of every rank is the
    multi-process one.  Returns the proofs of all ranks."""
    import threading
    from starky_bls12_381_b200.sharded import GpuBackend, ThreadGroup, prove_sharded
    plan = shard_plan(p.n_cols, p.log_n, p.rate_bits, world)
    shared = ThreadGroup.Shared(world)
    out, errs = [None] * world, []

    def run(rank):
        ctx = sb.Context(0)
        try:
            if airbin:
                ctx.air_load(stark_id, airbin)
            be = GpuBackend(ctx, p)
            c0, cg = plan.col_start[rank], plan.col_count[rank]
            out[rank] = prove_sharded(be, plan, rank, trace[c0:c0 + cg], pis, comm=ThreadGroup(shared, rank), fused=fused)
        except Exception as e:          # noqa: BLE001 -- a dead rank would hang the others at the barrier
            errs.append(e)
            shared.barrier.abort()
        finally:
            ctx.close()
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert not errs, errs
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("world,log_n", [(1, 6), (2, 6), (4, 7), (8, 8)])
def test_gpu_sharded_proof_of_a_valid_trace_equals_the_single_gpu_proof(world, log_n, tmp_path):
    import toy_air
    from helpers import to_oracle_params
    air = toy_air.limbs(str(tmp_path), 4)
    trace, pis = air["witness"](log_n)
    p = sb.Params(205, log_n, air["n_cols"], air["n_pis"], air["degree"], air["rate_bits"], 4, 2, 16, 84, 4, 5, 0, 0, 0)
    ctx = sb.Context(0)
    try:
        ctx.air_load(205, air["airbin"])
        want = ctx.prove(p, trace, pis)
    finally:
        ctx.close()
    proofs = _prove_sharded_threads(world, p, trace, pis, airbin=air["airbin"], stark_id=205)
    for pr in proofs:
        assert np.array_equal(pr.words, want.words)
    assert O.verify(air["flat"], to_oracle_params(p), proofs[0].words) == 0, O.err()


@pytest.mark.gpu
@pytest.mark.parametrize("name,world,log_n", [("miller_loop", 4, 6), ("pairing_precomp", 2, 5), ("ecc_agg", 8, 7)])
def test_gpu_sharded_proof_with_the_reference_constraint_programs(name, world, log_n):
    """Ragged column slices (97330 = 2 x 24333 + 2 x 24332), the real constraint programs, random traces."""
    from starky_bls12_381_b200 import airfiles
    info = sb.STARKS[name]
    airfiles.air_path(name, "airbin")
    p = sb.standard_params(info.stark_id, log_n, flags=sb.Flags.ALLOW_INVALID_TRACE)
    rng = np.random.default_rng(0xB2003000 + info.stark_id)
    trace = random_trace(rng, info.columns, log_n)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    ctx = sb.Context(0)
    try:
        want = ctx.prove(p, trace, pis)
    finally:
        ctx.close()
    for pr in _prove_sharded_threads(world, p, trace, pis):
        assert np.array_equal(pr.words, want.words)


# ---------------------------------------------------------------- the tail's small collectives under gloo (CPU)
def _tail_worker(rank, world, init_file, n_cols, log_n, rate_bits, q):
    from starky_bls12_381_b200.sharded import TorchGroup, combine_partials, gather_openings, gather_query_rows
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        plan = shard_plan(n_cols, log_n, rate_bits, world)
        comm = TorchGroup(world, rank)
        c0, cg = plan.col_start[rank], plan.col_count[rank]
        # openings: value of global column c is a function of c only
        cols = np.arange(c0, c0 + cg, dtype=np.uint64)
        mine = np.stack([np.stack([cols * np.uint64(3) + np.uint64(1), cols * np.uint64(5) + np.uint64(2)], axis=1),
                         np.stack([cols * np.uint64(7) + np.uint64(3), cols * np.uint64(11) + np.uint64(4)], axis=1)])
        op = gather_openings(comm, plan, rank, mine)
        # combine: partial sums near p so that the modular addition wraps
        n = 1 << log_n
        part = (np.arange(2 * n, dtype=np.uint64).reshape(n, 2) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(rank * 77)) % np.uint64(P)
        part[0, 0] = P - 1 - rank
        tot = combine_partials(comm, torch.from_numpy(part.view(np.int64)))
        # query rows: this rank's row block of a global [C][N] table whose entry is a function of (c, pos)
        R = plan.rows_per_rank
        cc, pp = np.meshgrid(np.arange(n_cols, dtype=np.int64), np.arange(rank * R, (rank + 1) * R, dtype=np.int64), indexing="ij")
        rows = torch.from_numpy(cc * 1000003 + pp)
        positions = [0, plan.n_lde - 1, R - 1, R % plan.n_lde, 5, plan.n_lde // 2 + 3, 5]
        qr = gather_query_rows(comm, plan, rank, rows, positions)
        q.put((rank, op, tot, qr.numpy().copy(), part))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_cols", [(2, 13), (4, 9), (4, 3)])
def test_tail_collectives_over_gloo(world, n_cols):
    log_n, rate_bits = 6, 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        procs = [ctx.Process(target=_tail_worker, args=(r, world, os.path.join(d, "rv"), n_cols, log_n, rate_bits, q)) for r in range(world)]
        for pr in procs:
            pr.start()
        res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
        for pr in procs:
            pr.join(timeout=60)
            assert pr.exitcode == 0
    cols = np.arange(n_cols, dtype=np.uint64)
    want_op = np.stack([np.stack([cols * np.uint64(3) + np.uint64(1), cols * np.uint64(5) + np.uint64(2)], axis=1),
                        np.stack([cols * np.uint64(7) + np.uint64(3), cols * np.uint64(11) + np.uint64(4)], axis=1)])
    want_tot = np.zeros_like(res[0][4])
    acc = [[0, 0] for _ in range(want_tot.shape[0])]
    for _, _, _, _, part in res:
        for i in range(part.shape[0]):
            for j in range(2):
                acc[i][j] = (acc[i][j] + int(part[i, j])) % P
    want_tot = np.array(acc, dtype=np.uint64)
    N = 1 << (log_n + rate_bits)
    positions = [0, N - 1, N // world - 1, (N // world) % N, 5, N // 2 + 3, 5]
    want_rows = np.array([[c * 1000003 + p for c in range(n_cols)] for p in positions], dtype=np.int64)
    for rank, op, tot, qr, _ in res:
        assert np.array_equal(op, want_op)
        assert np.array_equal(tot, want_tot)
        assert np.array_equal(qr, want_rows)


@pytest.mark.gpu
@pytest.mark.parametrize("name,world,log_n", [("miller_loop", 4, 6), ("ecc_agg", 8, 7), ("pairing_precomp", 2, 8)])
def test_gpu_sharded_proof_with_k1_storing_into_peer_row_buffers(name, world, log_n):
    """fused=True: K1 writes the LDE straight into the row buffers of the owning ranks (sb_lde_cols_peer_device; here the
    'peers' are buffers of the other host threads on the same GPU), no all-to-all: same proof."""
    from starky_bls12_381_b200 import airfiles
    info = sb.STARKS[name]
    airfiles.air_path(name, "airbin")
    p = sb.standard_params(info.stark_id, log_n, flags=sb.Flags.ALLOW_INVALID_TRACE)
    rng = np.random.default_rng(0xB2004000 + info.stark_id)
    trace = random_trace(rng, info.columns, log_n)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    ctx = sb.Context(0)
    try:
        want = ctx.prove(p, trace, pis)
    finally:
        ctx.close()
    for pr in _prove_sharded_threads(world, p, trace, pis, fused=True):
        assert np.array_equal(pr.words, want.words)
