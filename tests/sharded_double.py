"""CPU double of starky_bls12_381_b200.sharded.GpuBackend for the gloo tests: the same three stage functions computed
by the CPU oracle (TEST INFRASTRUCTURE: lives under tests/, the product never imports it)."""
import numpy as np
import torch

import oracle_lib as O
from helpers import pos_to_leaf


class OracleBackend:
    def lde_cols(self, plan, rank, local_trace):
        cg = plan.col_count[rank]
        p = O.make_params(log_n=plan.log_n, n_cols=cg, rate_bits=plan.rate_bits)
        leaves = O.lde_commit(p, np.ascontiguousarray(local_trace, dtype=np.uint64))["leaves"]     # [N][C_g], plonky2 leaf order
        lde_pos = np.ascontiguousarray(leaves[pos_to_leaf(plan.log_n, plan.rate_bits)].T)            # [C_g][N], device position order
        slabs = lde_pos.reshape(cg, plan.world, plan.rows_per_rank).transpose(1, 0, 2)
        return torch.from_numpy(np.ascontiguousarray(slabs).reshape(-1).view(np.int64))

    def hash_rows(self, plan, rows):
        r = rows.numpy().view(np.uint64)
        dig = np.stack([O.hash_or_noop(np.ascontiguousarray(r[:, j])) for j in range(plan.rows_per_rank)])
        return torch.from_numpy(dig.view(np.int64))

    def merkle_cap(self, plan, digests, cap_height=4):
        d = digests.numpy().view(np.uint64)
        level = np.empty_like(d)
        level[pos_to_leaf(plan.log_n, plan.rate_bits)] = d          # position order -> plonky2 leaf order
        while level.shape[0] > (1 << cap_height):
            level = np.stack([O.two_to_one(level[2 * i], level[2 * i + 1]) for i in range(level.shape[0] // 2)])
        return level
