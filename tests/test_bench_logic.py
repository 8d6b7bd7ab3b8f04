"""CPU-only checks of bench.py's host logic (no GPU, no library calls)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_full_set_assignment_covers_the_seven_proofs_and_balances():
    for world in (1, 2, 4, 8):
        a = bench.full_set_assignment(world)
        assert len(a) == world
        assert sorted(x for r in a for x in r) == sorted(bench.FULL_SET)
        loads = [sum(bench.FULL_SET_COST[x] for x in r) for r in a]
        # FinalExp is the critical path as soon as there are two GPUs
        assert max(loads) == (sum(bench.FULL_SET_COST[x] for x in bench.FULL_SET) if world == 1 else bench.FULL_SET_COST["final_exp"])
    assert bench.full_set_assignment(4)[0] == ["final_exp"]


def test_workload_tables_are_consistent():
    import starky_bls12_381_b200 as sb
    assert set(bench.WORKLOADS) == set(sb.STARKS) == set(bench.K_CONSTRAINTS)
    assert set(bench.FULL_SET) == set(sb.STARKS)


def test_full_set_plan_shards_final_exp_on_a_whole_box():
    rest = sorted(k for k in bench.FULL_SET if k != "final_exp")
    # default plan: FinalExp over every GPU, the six other proofs longest-first over the same GPUs (next to the shards)
    for world in (4, 8):
        fe, per = bench.full_set_plan(world)
        assert fe == list(range(world)) and len(per) == world
        assert sorted(x for r in per for x in r) == rest
        assert max(sum(bench.FULL_SET_COST[x] for x in r) for r in per) == bench.FULL_SET_COST["miller_loop"]
    # round-1 plan: FinalExp over half of the GPUs, the others share the rest
    fe, per = bench.full_set_plan(8, "half")
    assert fe == [0, 1, 2, 3] and per[:4] == [[], [], [], []]
    assert sorted(x for r in per for x in r) == rest
    assert max(sum(bench.FULL_SET_COST[x] for x in r) for r in per) == bench.FULL_SET_COST["miller_loop"]
    fe, per = bench.full_set_plan(4, "half")
    assert fe == [0, 1] and per[:2] == [[], []]
    assert sorted(x for r in per for x in r) == rest
    fe, per = bench.full_set_plan(2)
    assert fe == [] and per == bench.full_set_assignment(2)


def test_symmetric_memory_probe_plan_is_valid_at_every_world_size():
    """The probe that decides fused K1 vs all-to-all must never fail on its own plan: with a 64-position plan it was rejected
    by shard_plan at 4 and 8 ranks and the sharded legs fell back to the all-to-all without anyone asking them to."""
    import bench
    for world in (1, 2, 4, 8, 16, 32, 64):
        plan = bench.probe_plan(world)
        assert plan.world == world and plan.rows_per_rank >= 32 and sum(plan.col_count) == plan.n_cols
