"""Order-sensitive fingerprints of a constraint program: cs_refs, cs_class, class counts,
number of constraints touching next_values.  Used to cross-check the extraction against the survey's independent scan."""
from .rustsym import ADD, CONST, LOCAL, MUL, NEXT, PI, SUB

M = (1 << 61) - 1
W = {LOCAL: 1, NEXT: 1000003, PI: 1000000007}


def ref_sets(dag):
    sets = [None] * len(dag.nodes)
    empty = frozenset()
    for i, (op, a, b) in enumerate(dag.nodes):      # nodes are in topological order by construction
        if op == CONST:
            sets[i] = empty
        elif op in (LOCAL, NEXT, PI):
            sets[i] = frozenset(((op, a),))
        else:
            sa, sb = sets[a], sets[b]
            sets[i] = sa if sb <= sa else (sb if sa <= sb else sa | sb)
    return sets


def fingerprint(dag, constraints):
    sets = ref_sets(dag)
    cs_refs = cs_class = 0
    counts = {1: 0, 2: 0, 3: 0, 4: 0}
    uses_next = 0
    cols, pis = set(), set()
    for k, (cls, nid) in enumerate(constraints):
        r = 0
        nx = False
        for (t, i) in sets[nid]:
            r += (i + 1) * W[t]
            if t == NEXT: nx = True
            if t == PI: pis.add(i)
            else: cols.add(i)
        cs_refs = (cs_refs + (k + 1) * (r % M)) % M
        cs_class = (cs_class + (k + 1) * cls) % M
        counts[cls] += 1
        uses_next += nx
    return dict(K=len(constraints), plain=counts[1], transition=counts[2], first=counts[3], last=counts[4],
                uses_next=uses_next, cs_refs=cs_refs, cs_class=cs_class, distinct_cols=len(cols),
                max_col=max(cols) if cols else -1, distinct_pis=len(pis))
