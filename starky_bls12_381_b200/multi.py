"""Multi-GPU groups (include/starky_b200.h "multi-GPU groups", csrc/group.cu): thin ctypes callers.  The sharded proof --
peer stores over NVLink, NCCL collectives, the modular add of the combine partials -- lives in libstarkyb200; Python only
decides which GPUs form a group and, for one process per GPU, carries the 128-byte NCCL id from the first rank to the
others (here over torch.distributed, which the bench already uses for its barrier)."""
import ctypes as C
import threading

import numpy as np

from . import binding as B

PHASES = ("commit", "quotient", "openings", "combine", "query_rows")


def shard_columns(p, world, rank):
    """(first_col, n_cols_local, rows_per_rank) of `rank` in a `world`-rank group: sb_shard_columns."""
    c0, cg, rows = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = B.lib().sb_shard_columns(C.byref(p), world, rank, C.byref(c0), C.byref(cg), C.byref(rows))
    if rc:
        raise B.SbError(rc, B.lib().sb_last_error(None).decode())
    return c0.value, cg.value, rows.value


class Group:
    """One rank of a multi-GPU group (sb_group).  `member` is False on ranks outside a sub-group."""

    def __init__(self, ctx, handle, rank, world):
        self.ctx, self._g, self.rank, self.world, self.member = ctx, handle, rank, world, handle is not None
        self.fused_ok = True

    @classmethod
    def from_torch(cls, ctx, rank, world, local_rank, ranks=None):
        """One process per GPU (torchrun): NCCL communicator inside the library.  Collective over ALL ranks of the default
        torch process group; `ranks` (default: all) are the members of the new group, the others get a non-member object."""
        import torch
        import torch.distributed as dist
        ranks = list(range(world)) if ranks is None else list(ranks)
        ident = np.zeros(128, np.uint8)
        if rank == ranks[0]:
            rc = B.lib().sb_group_unique_id(ident.ctypes.data_as(C.c_void_p))
            if rc:
                raise B.SbError(rc, B.lib().sb_last_error(None).decode())
        t = torch.from_numpy(ident).to(torch.device("cuda", local_rank))
        if world > 1:
            dist.broadcast(t, ranks[0])
        ident = t.cpu().numpy().copy()
        if rank not in ranks:
            return cls(ctx, None, -1, len(ranks))
        h = C.c_void_p()
        rc = B.lib().sb_group_init_rank(ctx._h, ranks.index(rank), len(ranks), ident.ctypes.data_as(C.c_void_p), C.byref(h))
        if rc:
            raise B.SbError(rc, B.lib().sb_last_error(None).decode())
        return cls(ctx, h, ranks.index(rank), len(ranks))

    @classmethod
    def local(cls, contexts):
        """One process, one context per rank (several GPUs, or one GPU shared by all ranks in the tests): peer copies and
        host barriers, no NCCL.  Returns one Group per context; prove() must be called from one thread per rank."""
        n = len(contexts)
        hs = (C.c_void_p * n)(*[c._h for c in contexts])
        out = (C.c_void_p * n)()
        rc = B.lib().sb_group_init_local(hs, n, out)
        if rc:
            raise B.SbError(rc, B.lib().sb_last_error(None).decode())
        return [cls(contexts[r], C.c_void_p(out[r]), r, n) for r in range(n)]

    def column_slice(self, p):
        return shard_columns(p, self.world, max(self.rank, 0))[:2]

    def prove(self, p, local_trace, public_inputs, on_device=False, fused=True):
        """Collective.  local_trace: this rank's [n_cols_local][n] uint64 columns -- numpy array or raw pointer (host, or
        device with on_device=True)."""
        out = C.POINTER(B._CProof)()
        pis = np.ascontiguousarray(public_inputs, dtype=np.uint64)
        if isinstance(local_trace, np.ndarray):
            local_trace = np.ascontiguousarray(local_trace, dtype=np.uint64)
            self._keep = local_trace
        rc = B.lib().sb_group_prove(self._g, C.byref(p), B._ptr(local_trace), int(on_device), B._ptr(pis) if pis.size else None,
                                    0 if fused else 1, C.byref(out))
        self.ctx._check(rc)
        proof = B.Proof(out)
        proof.phase_ms = {k: round(float(B.lib().sb_group_phase_ms(self._g, k.encode())), 3) for k in PHASES}
        self.fused_ok = B.lib().sb_group_phase_ms(self._g, b"fused") > 0.5
        return proof

    def close(self):
        if self._g is not None:
            B.lib().sb_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prove_on_local_group(groups, p, slices, public_inputs, on_device=False, fused=True):
    """Runs one sharded proof on a Group.local() set: one host thread per rank, returns the list of proofs (all equal)."""
    out, errs = [None] * len(groups), [None] * len(groups)

    def run(r):
        try:
            out[r] = groups[r].prove(p, slices[r], public_inputs, on_device=on_device, fused=fused)
        except Exception as e:      # noqa: BLE001
            errs[r] = e
    ts = [threading.Thread(target=run, args=(r,)) for r in range(len(groups))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return out


def same_on_every_rank(words):
    """True if every rank of the default torch process group holds the same proof words (checksum + caps compared)."""
    import torch
    import torch.distributed as dist
    w = np.ascontiguousarray(words, dtype=np.uint64)
    digest = np.concatenate([w[:64], np.array([np.bitwise_xor.reduce(w), w.sum(dtype=np.uint64), w.size], np.uint64)])
    t = torch.from_numpy(digest.view(np.int64).copy()).cuda()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return True
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return all(bool(torch.equal(parts[0], x)) for x in parts)
