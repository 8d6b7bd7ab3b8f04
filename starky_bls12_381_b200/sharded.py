"""Trace commitment sharded over the GPUs of one box: the multi-GPU form of
PolynomialBatch::from_values as called by starky::prover::prove (aggregate_proof.rs:59,105,138,169,212).

    phase 1  column-sharded   rank g owns columns [c0_g, c0_g + C_g): K1 (iNTT + coset LDE) on its slice, written as
                              `world` slabs [b][c][N / world] so that every destination's slab is contiguous
    exchange all-to-all       slab b of every rank -> rank b (NCCL over NVLink; LDE bytes x (G-1)/G cross the fabric)
    phase 2  row-sharded      rank g holds all C columns of its N / world positions: K2 leaf sponge, row-local
    gather   all-gather       N x 32-byte digests (<= 1 MB); K3 tree to the cap, redundantly on every rank
    phase 2' row-sharded      K4 quotient values of the rank's positions (quotient_sharded): the alphas come from the
                              gathered cap (same transcript on every rank); the "next" row of a block's last position
                              lives on another rank when a block is shorter than one coset, so the ranks all-gather
                              their FIRST rows (C x 8 bytes each) and each picks its successor's; the 2 x N quotient
                              values (<= 0.5 MB) are all-gathered for the (small, unsharded) quotient commitment

The module is backend-agnostic plumbing (torch.distributed only): `backend` supplies the three kernels -- GpuBackend
(libstarkyb200 through the C ABI, device pointers of torch CUDA tensors) in production, an oracle-backed CPU double in
the gloo tests.  Tensors are int64 views of canonical u64 field elements.
"""
import dataclasses

import numpy as np


@dataclasses.dataclass(frozen=True)
class ShardPlan:
    n_cols: int
    log_n: int
    rate_bits: int
    world: int
    col_start: tuple        # first column of every rank
    col_count: tuple        # columns of every rank (ragged: 97330 = 8 * 12166 + 2)
    rows_per_rank: int      # LDE positions hashed by every rank

    @property
    def n_lde(self):
        return 1 << (self.log_n + self.rate_bits)

    def send_splits(self, rank):
        """elements sent by `rank` to every destination (its columns x the destination's positions)"""
        return [self.col_count[rank] * self.rows_per_rank] * self.world

    def recv_splits(self, rank):
        """elements received by `rank` from every source (the source's columns x this rank's positions)"""
        return [c * self.rows_per_rank for c in self.col_count]

    def a2a_bytes_out(self, rank):
        return 8 * self.col_count[rank] * self.rows_per_rank * (self.world - 1)


def shard_plan(n_cols, log_n, rate_bits, world):
    if world < 1 or world & (world - 1):
        raise ValueError("world size %d: the row blocks are power-of-two slices of the LDE domain" % world)
    n_lde = 1 << (log_n + rate_bits)
    if n_lde // world < 32:
        raise ValueError("fewer than 32 LDE positions per rank")
    base, extra = divmod(n_cols, world)
    counts = tuple(base + (1 if g < extra else 0) for g in range(world))
    starts = tuple(int(x) for x in np.concatenate([[0], np.cumsum(counts)[:-1]]))
    return ShardPlan(n_cols, log_n, rate_bits, world, starts, counts, n_lde // world)


def commit_sharded(backend, plan, rank, local_trace, group=None):
    """Runs the sharded commitment on this rank.  local_trace: this rank's columns, [C_rank][n].
    Returns dict(cap=[2^cap_height][4] uint64 numpy, digests=[N][4] tensor in device position order, rows=tensor
    [C][N/world] (this rank's row block of the LDE, all columns))."""
    import torch
    import torch.distributed as dist
    slabs = backend.lde_cols(plan, rank, local_trace)                       # [world][C_rank][rows] flattened
    if plan.world == 1:
        rows = slabs
    else:
        rows = torch.empty(plan.n_cols * plan.rows_per_rank, dtype=torch.int64, device=slabs.device)
        dist.all_to_all_single(rows, slabs, output_split_sizes=plan.recv_splits(rank),
                               input_split_sizes=plan.send_splits(rank), group=group)
    rows = rows.view(plan.n_cols, plan.rows_per_rank)
    dig = backend.hash_rows(plan, rows)                                     # [rows][4], position order
    if plan.world == 1:
        digests = dig
    else:
        parts = [torch.empty_like(dig) for _ in range(plan.world)]
        dist.all_gather(parts, dig, group=group)
        digests = torch.cat(parts, dim=0)
    cap = backend.merkle_cap(plan, digests)
    return dict(cap=cap, digests=digests, rows=rows)


def successor_block(plan, rank):
    """The rank whose first row is the `next` row of this rank's last position (position J*n + k, next = J*n + (k+1) mod n):
    the following block inside the coset, wrapping to the coset's first block."""
    blocks_per_coset = (1 << plan.log_n) // plan.rows_per_rank
    if blocks_per_coset <= 1:
        return None                      # a block holds whole cosets: every successor is local
    nxt = rank + 1
    return nxt if nxt % blocks_per_coset else nxt - blocks_per_coset


def quotient_sharded(backend, plan, rank, rows, cap, public_inputs, group=None):
    """Row-sharded quotient evaluation (this rank's share of starky::prover::compute_quotient_polys).
    rows: [C][N/world] from commit_sharded; cap: the trace cap every rank holds.  Returns dict(alphas, q = tensor
    [2][N] in device position order, gathered on every rank; q_local = this rank's [2][N/world])."""
    import torch
    import torch.distributed as dist
    alphas = backend.alphas(cap)
    halo = None
    nb = successor_block(plan, rank)
    if nb is not None:
        first = rows[:, 0].contiguous()
        parts = [torch.empty_like(first) for _ in range(plan.world)]
        dist.all_gather(parts, first, group=group)
        halo = parts[nb]
    q_local = backend.quotient_rows(plan, rank, rows, halo, public_inputs, alphas)          # [2][rows_per_rank]
    if plan.world == 1:
        q = q_local
    else:
        parts = [torch.empty_like(q_local) for _ in range(plan.world)]
        dist.all_gather(parts, q_local, group=group)
        q = torch.cat(parts, dim=1)
    return dict(alphas=alphas, q=q, q_local=q_local)


class GpuBackend:
    """The three kernels through the C ABI (sb_lde_cols_device / sb_hash_rows_device / sb_merkle_from_position_digests)."""

    def __init__(self, ctx, params):
        import torch
        from . import binding as B
        self.ctx, self.p, self.B, self.torch = ctx, params, B, torch
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _sync_torch(self):
        self.torch.cuda.current_stream().synchronize()      # torch's stream -> the ctx stream hand-off

    def lde_cols(self, plan, rank, local_trace):
        t = self.torch
        cg, n = plan.col_count[rank], 1 << plan.log_n
        if not isinstance(local_trace, t.Tensor):
            local_trace = t.from_numpy(np.ascontiguousarray(local_trace, dtype=np.uint64).view(np.int64))
        d_trace = local_trace.to(self.device, non_blocking=False).contiguous()
        assert d_trace.numel() == cg * n
        out = t.empty(cg * plan.n_lde, dtype=t.int64, device=self.device)
        self.coeffs = t.empty(cg * n, dtype=t.int64, device=self.device)   # stays column-sharded (openings, FRI combine)
        self._sync_torch()
        self.ctx._check(self.B.lib().sb_lde_cols_device(self.ctx._h, self.B.C.byref(self.p), d_trace.data_ptr(), cg, plan.world,
                                                      self.coeffs.data_ptr(), out.data_ptr()))
        self.ctx.synchronize()
        return out

    def hash_rows(self, plan, rows):
        t = self.torch
        dig = t.empty((plan.rows_per_rank, 4), dtype=t.int64, device=self.device)
        self._sync_torch()
        self.ctx._check(self.B.lib().sb_hash_rows_device(self.ctx._h, rows.data_ptr(), plan.n_cols, plan.rows_per_rank, dig.data_ptr()))
        self.ctx.synchronize()
        return dig

    def merkle_cap(self, plan, digests):
        cap = np.empty((1 << self.p.cap_height, 4), np.uint64)
        self._sync_torch()
        self.ctx._check(self.B.lib().sb_merkle_from_position_digests(self.ctx._h, self.B.C.byref(self.p), digests.data_ptr(),
                                                                   cap.ctypes.data_as(self.B.C.c_void_p)))
        return cap

    def alphas(self, cap):
        cap = np.ascontiguousarray(cap, dtype=np.uint64)
        out = np.zeros(self.p.num_challenges, np.uint64)
        rc = self.B.lib().sb_transcript_alphas(cap.ctypes.data_as(self.B.C.c_void_p), cap.shape[0], self.p.num_challenges,
                                               out.ctypes.data_as(self.B.C.c_void_p))
        if rc:
            raise self.B.SbError(rc, "sb_transcript_alphas")
        return out

    def quotient_rows(self, plan, rank, rows, halo, public_inputs, alphas):
        t = self.torch
        out = t.empty((2, plan.rows_per_rank), dtype=t.int64, device=self.device)
        pis = np.ascontiguousarray(public_inputs, dtype=np.uint64)
        al = np.ascontiguousarray(alphas, dtype=np.uint64)
        self._sync_torch()
        self.ctx._check(self.B.lib().sb_quotient_rows_device(
            self.ctx._h, self.B.C.byref(self.p), rows.data_ptr(), plan.rows_per_rank, rank,
            halo.data_ptr() if halo is not None else None, pis.ctypes.data_as(self.B.C.c_void_p) if pis.size else None,
            al.ctypes.data_as(self.B.C.c_void_p), out.data_ptr()))
        return out
