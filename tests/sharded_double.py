"""CPU double of starky_bls12_381_b200.sharded.GpuBackend for the gloo tests: the same three stage functions computed
by the CPU oracle (TEST INFRASTRUCTURE: lives under tests/, the product never imports it)."""
import numpy as np
import torch

import oracle_lib as O
from helpers import pos_to_leaf


class OracleBackend:
    """`air`, `params`, `full_trace`: only for quotient_rows (the double evaluates the whole quotient with the oracle and
    returns the rank's block; it also CHECKS that the halo row it was handed is the successor of the block's last row)."""

    def __init__(self, air=None, params=None, full_trace=None):
        self.air, self.params, self.full_trace = air, params, full_trace

    def alphas(self, cap):
        return O.challenger_run(np.ascontiguousarray(cap, dtype=np.uint64).reshape(-1), 2)

    def quotient_rows(self, plan, rank, rows, halo, public_inputs, alphas):
        from helpers import pos_to_natural
        n, R = 1 << plan.log_n, plan.rows_per_rank
        r = rows.numpy().view(np.uint64)
        last = rank * R + R - 1
        nxt = (last & ~(n - 1)) | ((last + 1) & (n - 1))
        if not (rank * R <= nxt < (rank + 1) * R):
            leaves = O.lde_commit(self.params, self.full_trace)["leaves"]
            lde_pos = leaves[pos_to_leaf(plan.log_n, plan.rate_bits)].T
            assert halo is not None and np.array_equal(halo.numpy().view(np.uint64), lde_pos[:, nxt]), "wrong halo row"
            assert np.array_equal(r[:, -1], lde_pos[:, last])
        q_nat = O.quotient_values(self.air, self.params, self.full_trace, public_inputs, alphas)     # [2][N], natural LDE index
        q_pos = q_nat[:, pos_to_natural(plan.log_n, plan.rate_bits)]
        return torch.from_numpy(np.ascontiguousarray(q_pos[:, rank * R:(rank + 1) * R]).view(np.int64))

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
