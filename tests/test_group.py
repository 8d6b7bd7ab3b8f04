"""Multi-GPU groups behind the C ABI (csrc/group.cu; include/starky_b200.h "multi-GPU groups").

CPU: the shard arithmetic of the library against the Python planner the gloo tests use.
GPU: sb_group_prove with the LocalComm transport -- `world` ranks, one sb_ctx and one host thread each, all on the one GPU
of the test box, peer stores and peer copies between the ranks' buffers -- must produce, on every rank, the proof of the
single-GPU sb_prove word for word, for valid traces (oracle verifier accepts) and for the reference's constraint programs
with ragged column slices; both with K1 storing into the owners' row buffers and with the all-to-all path.  No torch
collective is involved anywhere (the two-process NCCL transport is exercised by bench.py --gpus 2)."""
import numpy as np
import pytest

import oracle_lib as O
import starky_bls12_381_b200 as sb
from helpers import random_trace, to_oracle_params
from starky_bls12_381_b200 import multi
from starky_bls12_381_b200.sharded import shard_plan


def test_library_shard_arithmetic_equals_the_python_planner():
    for info in sb.STARKS.values():
        p = sb.standard_params(info.stark_id, info.num_rows.bit_length() - 1)
        for world in (1, 2, 4, 8):
            try:
                plan = shard_plan(info.columns, p.log_n, p.rate_bits, world)
            except ValueError:
                with pytest.raises(sb.SbError):
                    multi.shard_columns(p, world, 0)
                continue
            assert sum(plan.col_count) == info.columns
            for r in range(world):
                assert multi.shard_columns(p, world, r) == (plan.col_start[r], plan.col_count[r], plan.rows_per_rank)
    p = sb.standard_params(sb.StarkId.MILLER_LOOP, 10)
    with pytest.raises(sb.SbError):
        multi.shard_columns(p, 3, 0)                 # not a power of two
    with pytest.raises(sb.SbError):
        multi.shard_columns(p, 4, 4)                 # rank out of range


def _group_proofs(world, p, trace, pis, fused, airbin=None, stark_id=None, on_device=False):
    ctxs = [sb.Context(0) for _ in range(world)]
    groups = []
    try:
        if airbin:
            for c in ctxs:
                c.air_load(stark_id, airbin)
        groups = multi.Group.local(ctxs)
        slices = []
        for r in range(world):
            c0, cg = groups[r].column_slice(p)
            slices.append(np.ascontiguousarray(trace[c0:c0 + cg]))
        keep = []
        if on_device:
            import torch
            keep = [torch.from_numpy(s.view(np.int64)).cuda() for s in slices]
            slices = [t.data_ptr() if t.numel() else 0 for t in keep]
        proofs = multi.prove_on_local_group(groups, p, slices, pis, on_device=on_device, fused=fused)
        again = multi.prove_on_local_group(groups, p, slices, pis, on_device=on_device, fused=fused)   # buffers reused
        for a, b in zip(proofs, again):
            assert np.array_equal(a.words, b.words)
        assert all(g.fused_ok == (fused and world > 1) for g in groups)
        return proofs
    finally:
        for g in groups:
            g.close()
        for c in ctxs:
            c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,log_n,fused", [(1, 6, True), (2, 6, True), (2, 6, False), (4, 7, True), (8, 8, True), (8, 8, False)])
def test_group_proof_of_a_valid_trace_equals_the_single_gpu_proof(world, log_n, fused, tmp_path):
    import toy_air
    air = toy_air.limbs(str(tmp_path), 4)
    trace, pis = air["witness"](log_n)
    p = sb.Params(205, log_n, air["n_cols"], air["n_pis"], air["degree"], air["rate_bits"], 4, 2, 16, 84, 4, 5, 0, 0, 0)
    ctx = sb.Context(0)
    try:
        ctx.air_load(205, air["airbin"])
        want = ctx.prove(p, trace, pis)
    finally:
        ctx.close()
    proofs = _group_proofs(world, p, trace, pis, fused, airbin=air["airbin"], stark_id=205)
    for pr in proofs:
        assert np.array_equal(pr.words, want.words)
        assert set(pr.phase_ms) == set(multi.PHASES) and all(v >= 0 for v in pr.phase_ms.values())
    assert O.verify(air["flat"], to_oracle_params(p), proofs[0].words) == 0, O.err()


@pytest.mark.gpu
@pytest.mark.parametrize("name,world,log_n,fused,on_device", [
    ("miller_loop", 4, 6, True, False), ("miller_loop", 8, 10, True, True), ("pairing_precomp", 2, 5, False, False),
    ("ecc_agg", 8, 7, True, False), ("final_exp", 4, 5, True, True), ("final_exp", 8, 6, False, False)])
def test_group_proof_with_the_reference_constraint_programs(name, world, log_n, fused, on_device):
    """Ragged column slices (97330 = 2 x 24333 + 2 x 24332 on four ranks), the real constraint programs, random traces, row
    blocks shorter than one coset (halo rows from the successor rank): every rank's proof == the single-GPU proof."""
    info = sb.STARKS[name]
    p = sb.standard_params(info.stark_id, log_n, flags=sb.Flags.ALLOW_INVALID_TRACE)
    rng = np.random.default_rng(0xB2005000 + info.stark_id)
    trace = random_trace(rng, info.columns, log_n)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    ctx = sb.Context(0)
    try:
        want = ctx.prove(p, trace, pis)
    finally:
        ctx.close()
    for pr in _group_proofs(world, p, trace, pis, fused, on_device=on_device):
        assert np.array_equal(pr.words, want.words)


@pytest.mark.gpu
def test_one_ctx_over_several_devices_shards_inside_sb_prove():
    """sb_init(devices, n > 1): the caller sees ONE ctx; sb_prove cuts the host trace into column slices and proves on all
    of them (here the same GPU twice -- the test box has one).  Both host trace layouts, proof == single-GPU proof."""
    import ctypes
    info = sb.STARKS["miller_loop"]
    p = sb.standard_params(info.stark_id, 7, flags=sb.Flags.ALLOW_INVALID_TRACE)
    rng = np.random.default_rng(77)
    trace = random_trace(rng, info.columns, 7)
    pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
    one = sb.Context(0)
    many = sb.Context([0, 0, 0, 0])
    try:
        want = one.prove(p, trace, pis)
        got = many.prove(p, trace, pis)
        assert np.array_equal(got.words, want.words)
        cols = [np.array(trace[c], copy=True) for c in range(info.columns)]
        ptrs = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        got2 = many.prove(p, ctypes.addressof(ptrs), pis, sb.TraceLayout.COLS_U64_PTRS)
        assert np.array_equal(got2.words, want.words)
        assert many.kernel_launches() > 0
        with pytest.raises(sb.SbError):
            many.prove(p, None, pis, sb.TraceLayout.DEVICE_COLMAJOR_U64)
    finally:
        one.close()
        many.close()


@pytest.mark.gpu
def test_prove_batch_schedules_the_seven_proofs_and_returns_them_in_job_order():
    """sb_prove_batch: the job list of one BLS verification (aggregate_proof.rs:279-370 order), at reduced
    heights so that both job classes occur (32768 LDE positions = throughput-bound, the others latency-bound), on three
    contexts: every proof equals the one sb_prove gives alone; a bad job fails alone with its own code."""
    from starky_bls12_381_b200.binding import prove_batch
    names = ["pairing_precomp", "pairing_precomp", "miller_loop", "miller_loop", "fp12_mul", "final_exp", "ecc_agg"]
    logs = {"pairing_precomp": 6, "miller_loop": 5, "fp12_mul": 4, "final_exp": 6, "ecc_agg": 13}
    rng = np.random.default_rng(0xB2006000)
    jobs, want = [], []
    one = sb.Context(0)
    ctxs = [sb.Context(0) for _ in range(3)]
    try:
        for nm in names:
            info = sb.STARKS[nm]
            p = sb.standard_params(info.stark_id, logs[nm], flags=sb.Flags.ALLOW_INVALID_TRACE)
            trace = random_trace(rng, info.columns, logs[nm])
            pis = rng.integers(0, 1 << 32, info.public_inputs, dtype=np.uint64)
            jobs.append((p, trace, sb.TraceLayout.COLMAJOR_U64, pis))
            want.append(one.prove(p, trace, pis).words)
        bad = sb.standard_params(sb.StarkId.MILLER_LOOP, 5)
        bad.n_cols = 17                                            # does not match the constraint program
        jobs.append((bad, np.zeros((17, 32), np.uint64), sb.TraceLayout.COLMAJOR_U64, np.zeros(5064, np.uint64)))
        for contexts in (ctxs, ctxs[:1]):
            got = prove_batch(contexts, jobs)
            assert len(got) == len(jobs)
            for (res, ms), w in zip(got[:-1], want):
                assert np.array_equal(res.words, w) and ms > 0
            assert isinstance(got[-1][0], sb.SbError) and got[-1][0].code == -1
        # inside a batch every layout commits its trace column group by column group (capi.cu, ctx->yield_slabs): the same
        # jobs as row-major u32 rows (what the C++ witness generators write) give the same proofs
        rows = [(p, np.ascontiguousarray(t.T).astype(np.uint32), sb.TraceLayout.ROWMAJOR_U32, pis) for p, t, _, pis in jobs[:-1]]
        got = prove_batch(ctxs, rows)
        for (res, ms), w in zip(got, want):
            assert np.array_equal(res.words, w)
    finally:
        one.close()
        for c in ctxs:
            c.close()
