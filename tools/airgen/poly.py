"""Polynomial normal form of DAG nodes: {monomial (sorted tuple of leaf node ids): coefficient mod p}."""
from .rustsym import ADD, CONST, LOCAL, MUL, NEXT, P, PI, SUB


class Expander:
    def __init__(self, dag):
        self.dag = dag
        self.memo = {}

    def poly(self, nid):
        m = self.memo.get(nid)
        if m is not None:
            return m
        op, a, b = self.dag.nodes[nid]
        if op == CONST:
            v = self.dag.consts[a]
            r = {(): v} if v else {}
        elif op in (LOCAL, NEXT, PI):
            r = {(nid,): 1}
        elif op == ADD or op == SUB:
            r = dict(self.poly(a))
            for mono, c in self.poly(b).items():
                v = (r.get(mono, 0) + (c if op == ADD else P - c)) % P
                if v: r[mono] = v
                else: r.pop(mono, None)
        else:
            pa, pb = self.poly(a), self.poly(b)
            r = {}
            for ma, ca in pa.items():
                for mb, cb in pb.items():
                    mono = tuple(sorted(ma + mb))
                    v = (r.get(mono, 0) + ca * cb) % P
                    if v: r[mono] = v
                    else: r.pop(mono, None)
        self.memo[nid] = r
        return r

    def top_factors(self, nid):
        """Flatten the top-level product; constant-1 factors are dropped."""
        op, a, b = self.dag.nodes[nid]
        if op == MUL:
            return self.top_factors(a) + self.top_factors(b)
        if op == CONST and self.dag.consts[a] == 1:
            return []
        return [nid]

    def size(self, nid):
        # memo per instance: a process-wide table keyed by id(dag) handed the sizes of a collected DAG to a new one that
        # was allocated at the same address (a test that extracts several programs in one process failed once in a while)
        memo = self.__dict__.setdefault("_size_memo", {})
        if nid in memo: return memo[nid]
        op, a, b = self.dag.nodes[nid]
        s = 1 if op in (CONST, LOCAL, NEXT, PI) else 1 + self.size(a) + self.size(b)
        memo[nid] = s
        return s
