"""Trace commitment sharded over the GPUs of one box: the multi-GPU form of
PolynomialBatch::from_values as called by starky::prover::prove (aggregate_proof.rs:59,105,138,169,212).

    phase 1  column-sharded   rank g owns columns [c0_g, c0_g + C_g): K1 (iNTT + coset LDE) on its slice, written as
                              `world` slabs [b][c][N / world] so that every destination's slab is contiguous
    exchange all-to-all       slab b of every rank -> rank b (NCCL over NVLink; LDE bytes x (G-1)/G cross the fabric)
             or FUSED         (commit_sharded(fused=True)) K1 stores every LDE value straight into the row buffer of the
                              rank that owns its row block -- symmetric memory, peer pointers over NVLink
                              (sb_lde_cols_peer_device) -- so the transfer overlaps the transform and the all-to-all
                              pass disappears; two barriers replace it
    phase 2  row-sharded      rank g holds all C columns of its N / world positions: K2 leaf sponge, row-local
    gather   all-gather       N x 32-byte digests (<= 1 MB); K3 tree to the cap, redundantly on every rank
    phase 2' row-sharded      K4 quotient values of the rank's positions (quotient_sharded): the alphas come from the
                              gathered cap (same transcript on every rank); the "next" row of a block's last position
                              lives on another rank when a block is shorter than one coset, so the ranks all-gather
                              their FIRST rows (C x 8 bytes each) and each picks its successor's; the 2 x N quotient
                              values (<= 0.5 MB) are all-gathered for the (small, unsharded) quotient commitment
    tail     prove_sharded    every rank runs the library's proof orchestration (sb_prove_sharded: same transcript, same
                              proof everywhere); the five distributed steps are hooks implemented here: the commitment
                              and quotient above, the openings (column-sharded coefficient slices, all-gather of
                              2 C/G extension values), the FRI batch combine (per-rank partial sums with the rank's
                              alpha-power offset, all-gather of n extension values, added mod p) and the query rows
                              (84 x C values, summed over the ranks that own them)

THE PRODUCT PATH IS NOT HERE: since round 2 the sharded proof runs inside libstarkyb200 (csrc/group.cu -- NCCL or
peer copies, CUDA IPC row buffers, modular add of the combine partials on the device; thin ctypes caller: multi.py).
This module is the executable specification of that exchange pattern, kept because it runs WITHOUT a GPU: with the
oracle-backed double of the three kernels it is what the world-size-2/4/8 gloo tests on CPU exercise (shard plan, halo
successor, tail collectives), and with GpuBackend + ThreadGroup (ranks emulated by host threads on one GPU) it is the
test harness of the per-rank stage entry points of the C ABI (sb_lde_cols_device, sb_hash_rows_device, ...).
Tensors are int64 views of canonical u64 field elements.
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


def commit_sharded(backend, plan, rank, local_trace, group=None, comm=None, fused=False):
    """Runs the sharded commitment on this rank.  local_trace: this rank's columns, [C_rank][n].
    Returns dict(cap=[2^cap_height][4] uint64 numpy, digests=[N][4] tensor in device position order, rows=tensor
    [C][N/world] (this rank's row block of the LDE, all columns)).  fused: K1 writes into the peers' row buffers
    (no all-to-all); needs a GpuBackend and a comm with symmetric_rows()."""
    import torch
    comm = comm or TorchGroup(plan.world, rank, group)
    if fused and plan.world > 1:
        sym = getattr(backend, "_sym", None)
        if sym is None or sym[0] != (plan.n_cols, plan.rows_per_rank, plan.world):
            sym = backend._sym = ((plan.n_cols, plan.rows_per_rank, plan.world),) + tuple(comm.symmetric_rows(plan, backend.device))
        _, rows, peer_ptrs, barrier = sym
        barrier()                                   # every rank is done reading its rows of the previous proof
        backend.lde_cols_peer(plan, rank, local_trace, peer_ptrs)
        barrier()                                   # every rank has written its columns into my rows
    else:
        slabs = backend.lde_cols(plan, rank, local_trace)                   # [world][C_rank][rows] flattened
        rows = comm.all_to_all_rows(slabs, plan)
    rows = rows.view(plan.n_cols, plan.rows_per_rank)
    dig = backend.hash_rows(plan, rows)                                     # [rows][4], position order
    digests = torch.cat(comm.all_gather(dig), dim=0)
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


def quotient_sharded(backend, plan, rank, rows, cap, public_inputs, group=None, comm=None):
    """Row-sharded quotient evaluation (this rank's share of starky::prover::compute_quotient_polys).
    rows: [C][N/world] from commit_sharded; cap: the trace cap every rank holds.  Returns dict(alphas, q = tensor
    [2][N] in device position order, gathered on every rank; q_local = this rank's [2][N/world])."""
    import torch
    comm = comm or TorchGroup(plan.world, rank, group)
    alphas = backend.alphas(cap)
    halo = None
    nb = successor_block(plan, rank)
    if nb is not None:
        halo = comm.all_gather(rows[:, 0].contiguous())[nb]
    q_local = backend.quotient_rows(plan, rank, rows, halo, public_inputs, alphas)          # [2][rows_per_rank]
    q = torch.cat(comm.all_gather(q_local), dim=1)
    return dict(alphas=alphas, q=q, q_local=q_local)


P = 0xFFFFFFFF00000001


def _addmod(a, b):
    """(a + b) mod p on canonical uint64 numpy arrays."""
    s = a + b
    s = np.where(s < a, s + np.uint64(0xFFFFFFFF), s)          # wrapped: 2^64 = 2^32 - 1 (mod p); cannot wrap twice
    return np.where(s >= np.uint64(P), s - np.uint64(P), s)


class TorchGroup:
    """The collectives of prove_sharded over torch.distributed (NCCL on the GPUs of one box, gloo in the CPU tests)."""

    def __init__(self, world, rank, group=None):
        self.world, self.rank, self.group = world, rank, group

    def all_gather(self, t):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return [t]
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous(), group=self.group)
        return parts

    def sum_int64(self, t):
        """Plain integer sum (used where exactly one rank contributes a non-zero value per element)."""
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_to_all_rows(self, slabs, plan):
        """slab b of every rank -> rank b: [world][C_rank][rows] on every rank -> [C][rows] on every rank."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return slabs
        rows = torch.empty(plan.n_cols * plan.rows_per_rank, dtype=torch.int64, device=slabs.device)
        dist.all_to_all_single(rows, slabs, output_split_sizes=plan.recv_splits(self.rank),
                               input_split_sizes=plan.send_splits(self.rank), group=self.group)
        return rows


class ThreadGroup:
    """The same collectives between `world` host threads of ONE process (one sb_ctx per thread, all on one GPU): the
    single-GPU test double of the NCCL group -- the control flow of every rank is exactly the multi-process one."""

    class Shared:
        def __init__(self, world):
            import threading
            self.world, self.barrier, self.box = world, threading.Barrier(world), [None] * world

    def __init__(self, shared, rank):
        self.sh, self.world, self.rank, self.group = shared, shared.world, rank, None

    def _exchange(self, t):
        import torch
        torch.cuda.synchronize()
        self.sh.box[self.rank] = t
        self.sh.barrier.wait()
        parts = list(self.sh.box)
        self.sh.barrier.wait()
        return parts

    def all_gather(self, t):
        return [x.clone() for x in self._exchange(t.contiguous())]

    def sum_int64(self, t):
        parts = self._exchange(t)
        out = parts[0].clone()
        for x in parts[1:]:
            out += x
        return out

    def symmetric_rows(self, plan, device):
        import torch
        rows = torch.empty(plan.n_cols * plan.rows_per_rank, dtype=torch.int64, device=device)
        ptrs = [int(t.data_ptr()) for t in self._exchange(rows)]
        self._keep = rows

        def barrier():
            torch.cuda.synchronize()
            self.sh.barrier.wait()
        return rows, ptrs, barrier

    def all_to_all_rows(self, slabs, plan):
        import torch
        parts = self._exchange(slabs)
        mine = [parts[g].view(self.world, plan.col_count[g], plan.rows_per_rank)[self.rank].reshape(-1) for g in range(self.world)]
        return torch.cat(mine)


# ---- the three "small collectives" of the tail as plain data movement (tested under gloo on CPU) ----
def gather_openings(comm, plan, rank, mine, device=None):
    """mine: uint64 [2][C_rank][2] = (P_c(zeta), P_c(g zeta)) of this rank's columns.  Returns uint64 [2][C][2] on every
    rank (ragged column counts: every slice is padded to the widest before the all-gather, then cut)."""
    import torch
    cg, width = plan.col_count[rank], max(plan.col_count)
    pad = np.zeros((2, width, 2), np.uint64)
    pad[:, :cg] = mine[:, :cg]
    t = torch.from_numpy(pad.view(np.int64))
    parts = comm.all_gather(t.to(device) if device is not None else t)
    out = np.empty((2, plan.n_cols, 2), np.uint64)
    for g, part in enumerate(parts):
        a = part.cpu().numpy().view(np.uint64)
        out[:, plan.col_start[g]:plan.col_start[g] + plan.col_count[g]] = a[:, :plan.col_count[g]]
    return out


def combine_partials(comm, part):
    """part: int64 tensor [n][2], this rank's sum_c alpha^c coeffs_c over its columns (canonical field elements).
    Returns the sum over all ranks mod p as uint64 numpy [n][2] (NCCL has no modular reduction: all-gather, add here)."""
    total = None
    for t in comm.all_gather(part):
        a = t.cpu().numpy().view(np.uint64)
        total = a.copy() if total is None else _addmod(total, a)
    return total


def gather_query_rows(comm, plan, rank, rows, positions):
    """rows: int64 tensor [C][N/world] (this rank's row block); positions: device LDE positions of the queries.
    Returns int64 tensor [len(positions)][C]: every row comes from the rank that owns it, the others contribute zeros."""
    import torch
    pos = np.asarray(positions, np.int64)
    R = plan.rows_per_rank
    own = (pos // R) == rank
    out = torch.zeros((len(pos), plan.n_cols), dtype=torch.int64, device=rows.device)
    if own.any():
        idx = torch.from_numpy(pos[own] - rank * R).to(rows.device)
        out[torch.from_numpy(np.nonzero(own)[0]).to(rows.device)] = rows[:, idx].t()
    return comm.sum_int64(out).contiguous()


def prove_sharded(backend, plan, rank, local_trace, public_inputs, comm=None, fused=False):
    """One proof with the trace sharded over plan.world GPUs: every rank calls this with its column slice
    and gets the same proof.  backend: GpuBackend of this rank; comm: TorchGroup (default: the default process group)."""
    import torch
    from . import binding as B
    C = B.C
    comm = comm or TorchGroup(plan.world, rank)
    p, ctx, lib = backend.p, backend.ctx, B.lib()
    n = 1 << plan.log_n
    c0, cg = plan.col_start[rank], plan.col_count[rank]
    pis = np.ascontiguousarray(public_inputs, dtype=np.uint64)
    st = {}

    import time
    st["hook_ms"] = {}

    def guard(fn):
        def run(*a):
            try:
                t0 = time.perf_counter()
                fn(*a)
                st["hook_ms"][fn.__name__] = 1e3 * (time.perf_counter() - t0)
                return 0
            except B.SbError as e:
                st["error"] = e
                return e.code if e.code else -1
            except Exception as e:                      # never unwind through the C frames
                st["error"] = e
                return -1
        return run

    def h_commit(_user, cap_out):
        com = commit_sharded(backend, plan, rank, local_trace, comm=comm, fused=fused)
        st["rows"], st["cap"] = com["rows"], com["cap"]
        flat = np.ascontiguousarray(com["cap"], dtype=np.uint64).reshape(-1)
        C.memmove(cap_out, flat.ctypes.data, flat.nbytes)

    def h_quotient(_user, alphas, d_q_out):
        al = np.array([alphas[0], alphas[1]], np.uint64)
        halo, nb = None, successor_block(plan, rank)
        if nb is not None:
            halo = comm.all_gather(st["rows"][:, 0].contiguous())[nb]
        q_local = backend.quotient_rows(plan, rank, st["rows"], halo, pis, al)
        q = torch.cat(comm.all_gather(q_local), dim=1).contiguous()                    # [2][N], position order
        backend._sync_torch()
        ctx._check(lib.sb_memcpy_device(ctx._h, d_q_out, q.data_ptr(), 16 * plan.n_lde))

    def h_openings(_user, zeta, zeta_next, local_out, next_out):
        z = np.array([zeta[0], zeta[1]], np.uint64)
        zn = np.array([zeta_next[0], zeta_next[1]], np.uint64)
        mine = np.zeros((2, max(cg, 1), 2), np.uint64)
        if cg:
            ctx._check(lib.sb_openings_cols_device(ctx._h, C.byref(p), backend.coeffs.data_ptr(), cg, z.ctypes.data, zn.ctypes.data,
                                                   mine[0].ctypes.data, mine[1].ctypes.data))
        allv = gather_openings(comm, plan, rank, mine, backend.device)
        C.memmove(local_out, np.ascontiguousarray(allv[0]).ctypes.data, 16 * plan.n_cols)
        C.memmove(next_out, np.ascontiguousarray(allv[1]).ctypes.data, 16 * plan.n_cols)

    def h_combine(_user, alpha, d_out):
        al = np.array([alpha[0], alpha[1]], np.uint64)
        part = torch.zeros((n, 2), dtype=torch.int64, device=backend.device)
        backend._sync_torch()
        if cg:
            ctx._check(lib.sb_combine_cols_device(ctx._h, C.byref(p), backend.coeffs.data_ptr(), cg, al.ctypes.data, c0, part.data_ptr()))
        total = combine_partials(comm, part)
        dev = torch.from_numpy(total.view(np.int64)).to(backend.device)
        backend._sync_torch()
        ctx._check(lib.sb_memcpy_device(ctx._h, d_out, dev.data_ptr(), 16 * n))

    def h_query_rows(_user, positions, count, d_rows_out):
        rows = gather_query_rows(comm, plan, rank, st["rows"], [positions[i] for i in range(count)])
        backend._sync_torch()
        ctx._check(lib.sb_memcpy_device(ctx._h, d_rows_out, rows.data_ptr(), 8 * count * plan.n_cols))

    hooks = B.ShardHooks(None, B.HOOK_COMMIT(guard(h_commit)), B.HOOK_QUOTIENT(guard(h_quotient)),
                         B.HOOK_OPENINGS(guard(h_openings)), B.HOOK_COMBINE(guard(h_combine)),
                         B.HOOK_QUERY_ROWS(guard(h_query_rows)))
    out = C.POINTER(B._CProof)()
    rc = lib.sb_prove_sharded(ctx._h, C.byref(p), C.byref(hooks), pis.ctypes.data if pis.size else None, C.byref(out))
    if rc:
        if isinstance(st.get("error"), Exception) and not isinstance(st["error"], B.SbError):
            raise st["error"]
        ctx._check(rc)
    proof = B.Proof(out)
    proof.hook_ms = st["hook_ms"]          # wall milliseconds spent in each collective hook (includes waiting for peers)
    return proof


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

    def lde_cols_peer(self, plan, rank, local_trace, peer_ptrs):
        t = self.torch
        cg, n = plan.col_count[rank], 1 << plan.log_n
        if not isinstance(local_trace, t.Tensor):
            local_trace = t.from_numpy(np.ascontiguousarray(local_trace, dtype=np.uint64).view(np.int64))
        d_trace = local_trace.to(self.device, non_blocking=False).contiguous()
        assert d_trace.numel() == cg * n and len(peer_ptrs) == plan.world
        self.coeffs = t.empty(cg * n, dtype=t.int64, device=self.device)
        ptrs = np.array(peer_ptrs, dtype=np.uint64)
        self._sync_torch()
        self.ctx._check(self.B.lib().sb_lde_cols_peer_device(self.ctx._h, self.B.C.byref(self.p), d_trace.data_ptr(), cg, plan.world,
                                                           plan.col_start[rank], self.coeffs.data_ptr(), ptrs.ctypes.data))

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
