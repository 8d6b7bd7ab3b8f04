"""A small symbolic interpreter for the subset of Rust in which the reference writes its AIR constraints.

It plays the role a tracing `PackedField` type would play in a Rust build: run `Stark::eval_packed_generic`
(fp12_mul.rs:58, calc_pairing_precomp.rs:376, miller_loop.rs:644, final_exponentiate.rs:907, ecc_aggregate.rs:92
and every `add_*_constraints` gadget they call in fp.rs / fp2.rs / fp6.rs / fp12.rs / g1.rs) with symbolic
`local_values` / `next_values` / `public_inputs`, and record, in emission order, each polynomial handed to
`ConstraintConsumer::{constraint, constraint_transition, constraint_first_row, constraint_last_row}` as a node of a
hash-consed expression DAG.  Nothing from the reference is copied into this repository: the tool reads
/root/reference at generation time and writes only the derived constraint program (see tools/airgen/__main__.py).

Supported: consts with per-file shadowing and glob imports, fn calls, let / tuple patterns / mut / += , for over ranges,
if / else (statement and expression), closures, ranges, arrays, Option, `as`, method chains used by the gadgets
(.iter().map().collect(), .fold, .unwrap_or, .try_into().unwrap(), .concat(), ...), `bit_decomp_32!`, BigUint as int.
"""
import os
import re
import sys

sys.setrecursionlimit(20000)

P = 0xFFFFFFFF00000001

# ------------------------------------------------------------------------------------------------------------------
# expression DAG
# ------------------------------------------------------------------------------------------------------------------
CONST, LOCAL, NEXT, PI, ADD, SUB, MUL = range(7)


class Dag:
    def __init__(self):
        self.nodes = []          # (op, a, b)
        self.index = {}
        self.consts = []
        self.const_index = {}

    def _mk(self, op, a, b=0):
        key = (op, a, b)
        i = self.index.get(key)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(key)
            self.index[key] = i
        return i

    def const(self, v):
        v %= P
        ci = self.const_index.get(v)
        if ci is None:
            ci = len(self.consts)
            self.consts.append(v)
            self.const_index[v] = ci
        return self._mk(CONST, ci)

    def leaf(self, kind, idx):
        return self._mk(kind, idx)

    def binop(self, op, a, b):
        return self._mk(op, a, b)


class Sym:
    """A field-valued expression (P or FE in the reference's generics)."""
    __slots__ = ("dag", "id")

    def __init__(self, dag, nid):
        self.dag = dag
        self.id = nid

    def _coerce(self, o):
        if isinstance(o, Sym):
            return o
        if isinstance(o, int):
            return Sym(self.dag, self.dag.const(o))
        raise TypeError("cannot combine Sym with %r" % (o,))

    def __add__(self, o): return Sym(self.dag, self.dag.binop(ADD, self.id, self._coerce(o).id))
    def __radd__(self, o): return self._coerce(o) + self
    def __sub__(self, o): return Sym(self.dag, self.dag.binop(SUB, self.id, self._coerce(o).id))
    def __rsub__(self, o): return self._coerce(o) - self
    def __mul__(self, o): return Sym(self.dag, self.dag.binop(MUL, self.id, self._coerce(o).id))
    def __rmul__(self, o): return self._coerce(o) * self
    def __neg__(self): return Sym(self.dag, self.dag.const(0)) - self


class ColVec:
    """local_values / next_values / public_inputs."""

    def __init__(self, dag, kind, length):
        self.dag, self.kind, self.length = dag, kind, length

    def __getitem__(self, i):
        if isinstance(i, range):
            return [self[j] for j in i]
        if not (0 <= i < self.length):
            raise IndexError("%s index %d out of range %d" % ("LNP"[self.kind - 1], i, self.length))
        return Sym(self.dag, self.dag.leaf(self.kind, i))

    def __len__(self):
        return self.length


# ------------------------------------------------------------------------------------------------------------------
# tokenizer
# ------------------------------------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<num>0x[0-9a-fA-F_]+(?:u8|u16|u32|u64|u128|usize|i32|i64)?|\d[\d_]*(?:u8|u16|u32|u64|u128|usize|i32|i64)?)
  | (?P<life>'[A-Za-z_]\w*(?!'))
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>\.\.=|<<=|>>=|::|\.\.|->|=>|==|!=|<=|>=|&&|\|\||\+=|-=|\*=|/=|%=|<<|>>|[-+*/%=<>!&|^.,;:(){}\[\]#?$@])
""", re.X)


def strip_comments(src):
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if c == '"':
            j = i + 1
            while j < n and src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            out.append(src[i:j + 1]); i = j + 1
        elif src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif src.startswith("/*", i):
            depth, j = 1, i + 2
            while j < n and depth:
                if src.startswith("/*", j): depth += 1; j += 2
                elif src.startswith("*/", j): depth -= 1; j += 2
                else: j += 1
            out.append(" "); i = j
        else:
            out.append(c); i += 1
    return "".join(out)


def tokenize(src):
    toks, pos = [], 0
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError("cannot tokenize at %r" % src[pos:pos + 40])
        pos = m.end()
        k = m.lastgroup
        if k in ("ws", "life"):
            continue
        toks.append((k, m.group(k)))
    toks.append(("eof", ""))
    return toks


# ------------------------------------------------------------------------------------------------------------------
# parser (expressions + statements of fn bodies)
# ------------------------------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0): return self.t[self.i + k]
    def at(self, v, k=0): return self.t[self.i + k][1] == v and self.t[self.i + k][0] in ("op", "id")
    def next(self):
        tok = self.t[self.i]; self.i += 1; return tok

    def expect(self, v):
        tok = self.next()
        if tok[1] != v:
            raise SyntaxError("expected %r got %r near %s" % (v, tok[1], " ".join(x[1] for x in self.t[max(0, self.i - 12):self.i + 6])))
        return tok

    def accept(self, v):
        if self.at(v):
            self.i += 1
            return True
        return False

    # ---- skipping helpers ----
    def skip_angle(self):
        """at '<': skip a balanced generic-argument list."""
        assert self.at("<")
        depth = 0
        while True:
            tok = self.next()
            if tok[1] == "<": depth += 1
            elif tok[1] == "<<": depth += 2
            elif tok[1] == ">": depth -= 1
            elif tok[1] == ">>": depth -= 2
            elif tok[1] == "->": pass
            if depth <= 0: return

    def skip_type(self, stops):
        depth = 0
        while True:
            tok = self.peek()
            if tok[0] == "eof": return
            if depth == 0 and tok[1] in stops and tok[0] == "op": return
            if tok[1] in ("<", "(", "["): depth += 1
            elif tok[1] == "<<": depth += 2
            elif tok[1] in (">", ")", "]"): depth -= 1
            elif tok[1] == ">>": depth -= 2
            self.i += 1

    # ---- statements ----
    def block(self):
        self.expect("{")
        stmts = []
        while not self.at("}"):
            if self.accept(";"):
                continue
            stmts.append(self.stmt())
        self.expect("}")
        return ("block", stmts)

    def pattern(self):
        if self.accept("("):
            items = []
            while not self.at(")"):
                items.append(self.pattern())
                if not self.accept(","): break
            self.expect(")")
            return ("ptuple", items)
        self.accept("mut"); self.accept("ref"); self.accept("&")
        self.accept("mut")
        name = self.next()
        assert name[0] == "id", name
        return ("pname", name[1])

    def stmt(self):
        if self.at("let"):
            self.next()
            pat = self.pattern()
            if self.accept(":"):
                self.skip_type(("=", ";"))
            val = None
            if self.accept("="):
                val = self.expr()
            self.expect(";")
            return ("let", pat, val)
        if self.at("for"):
            self.next()
            pat = self.pattern()
            self.expect("in")
            it = self.expr(no_struct=True)
            body = self.block()
            return ("for", pat, it, body)
        if self.at("while"):
            self.next()
            cond = self.expr(no_struct=True)
            body = self.block()
            return ("while", cond, body)
        if self.at("return"):
            self.next()
            val = None if self.at(";") else self.expr()
            self.accept(";")
            return ("return", val)
        e = self.expr()
        if self.peek()[1] in ("=", "+=", "-=", "*=") and self.peek()[0] == "op":
            op = self.next()[1]
            rhs = self.expr()
            self.accept(";")
            return ("assign", op, e, rhs)
        if self.accept(";"):
            return ("expr", e, True)
        return ("expr", e, False)     # tail expression (or block-like statement)

    # ---- expressions ----
    BIN = [("||",), ("&&",), ("==", "!=", "<", ">", "<=", ">="), ("|",), ("^",), ("&",), ("<<", ">>"), ("+", "-"),
           ("*", "/", "%")]

    def expr(self, no_struct=False):
        lhs = self.binary(0)
        if self.peek()[0] == "op" and self.peek()[1] in ("..", "..="):
            op = self.next()[1]
            if self.peek()[1] in (")", "]", "}", ",", ";") or self.at("{"):
                rhs = None
            else:
                rhs = self.binary(0)
            return ("range", lhs, rhs, op == "..=")
        return lhs

    def binary(self, level):
        if level == len(self.BIN):
            return self.cast()
        lhs = self.binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self.BIN[level]:
            if self.peek()[1] == "|" and level == 3 and False:
                break
            op = self.next()[1]
            rhs = self.binary(level + 1)
            lhs = ("bin", op, lhs, rhs)
        return lhs

    def cast(self):
        e = self.unary()
        while self.at("as"):
            self.next()
            ty = self.next()[1]
            e = ("cast", e, ty)
        return e

    def unary(self):
        if self.peek()[0] == "op" and self.peek()[1] in ("-", "!", "&", "*"):
            op = self.next()[1]
            self.accept("mut") if op == "&" else None
            e = self.unary()
            if op in ("&", "*"):
                return e
            return ("un", op, e)
        if self.at("&&"):
            self.next()
            return self.unary()
        return self.postfix()

    def args(self):
        self.expect("(")
        out = []
        while not self.at(")"):
            out.append(self.expr())
            if not self.accept(","): break
        self.expect(")")
        return out

    def postfix(self):
        e = self.primary()
        while True:
            if self.at("("):
                e = ("call", e, self.args())
            elif self.at("["):
                self.next()
                idx = self.expr()
                self.expect("]")
                e = ("index", e, idx)
            elif self.at("?"):
                self.next()
            elif self.at("."):
                nxt = self.peek(1)
                if nxt[0] == "num":
                    self.next(); self.next()
                    e = ("field", e, int(nxt[1]))
                elif nxt[0] == "id":
                    self.next(); self.next()
                    name = nxt[1]
                    if self.at("::"):
                        self.next(); self.skip_angle()
                    if self.at("("):
                        e = ("method", e, name, self.args())
                    else:
                        e = ("field", e, name)
                else:
                    break
            else:
                break
        return e

    def primary(self):
        k, v = self.peek()
        if k == "num":
            self.next()
            m = re.match(r"(0x[0-9a-fA-F_]+|\d[\d_]*)", v)
            return ("lit", int(m.group(1).replace("_", ""), 0))
        if k == "str":
            self.next()
            return ("lit", v[1:-1])
        if v == "(" and k == "op":
            self.next()
            items = []
            trailing = False
            while not self.at(")"):
                items.append(self.expr())
                trailing = self.accept(",")
                if not trailing: break
            self.expect(")")
            if len(items) == 1 and not trailing:
                return items[0]
            return ("tuple", items)
        if v == "[" and k == "op":
            self.next()
            items = []
            if not self.at("]"):
                first = self.expr()
                if self.accept(";"):
                    cnt = self.expr()
                    self.expect("]")
                    return ("repeat", first, cnt)
                items.append(first)
                while self.accept(","):
                    if self.at("]"): break
                    items.append(self.expr())
            self.expect("]")
            return ("array", items)
        if v == "{" and k == "op":
            return self.block()
        if v == "if":
            return self.if_expr()
        if v in ("|", "||") and k == "op":
            return self.closure()
        if v == "move":
            self.next()
            return self.closure()
        if v == "<" and k == "op":      # <T>::NAME  (macro bodies)
            self.skip_angle()
            segs = ["<T>"]
            while self.accept("::"):
                segs.append(self.next()[1])
            return ("path", segs)
        if k == "id":
            self.next()
            segs = [v]
            while self.at("::"):
                self.next()
                if self.at("<"):
                    self.skip_angle()
                else:
                    segs.append(self.next()[1])
            if self.at("!"):            # macro invocation
                self.next()
                close = {"(": ")", "[": "]", "{": "}"}[self.peek()[1]]
                self.next()
                items = []
                while not self.at(close):
                    items.append(self.expr())
                    if self.accept(";"):
                        cnt = self.expr()
                        self.expect(close)
                        return ("macro", segs[-1], [("repeat", items[0], cnt)])
                    if not self.accept(","): break
                self.expect(close)
                return ("macro", segs[-1], items)
            return ("path", segs)
        raise SyntaxError("unexpected token %r near %s" % (v, " ".join(x[1] for x in self.t[max(0, self.i - 10):self.i + 8])))

    def if_expr(self):
        self.expect("if")
        cond = self.expr(no_struct=True)
        then = self.block()
        other = None
        if self.accept("else"):
            other = self.if_expr() if self.at("if") else self.block()
        return ("if", cond, then, other)

    def closure(self):
        params = []
        if self.accept("||"):
            pass
        else:
            self.expect("|")
            while not self.at("|"):
                params.append(self.pattern())
                if self.accept(":"):
                    self.skip_type((",", "|"))
                if not self.accept(","): break
            self.expect("|")
        body = self.expr()
        return ("closure", params, body)


# ------------------------------------------------------------------------------------------------------------------
# crate model: files, consts, fns
# ------------------------------------------------------------------------------------------------------------------
class RFile:
    def __init__(self, name, src):
        self.name = name
        self.src = strip_comments(src)
        self.const_src = {}
        self.const_val = {}
        self.fn_src = {}       # name -> (params, body_tokens_span) lazily parsed
        self.fn_ast = {}
        self.globs, self.explicit = [], {}
        self._scan()

    def _scan(self):
        s = self.src
        for m in re.finditer(r"\bconst\s+([A-Z_][A-Z0-9_]*)\s*:\s*[^=;]+=\s*([^;]+);", s):
            self.const_src.setdefault(m.group(1), m.group(2))
        for m in re.finditer(r"\buse\s+crate::([^;]+);", s):
            self._use(m.group(1).strip())
        for m in re.finditer(r"\bfn\s+(\w+)", s):
            name = m.group(1)
            if name in self.fn_src:
                continue
            j = m.end()
            if j < len(s) and s[j] == "<":      # generics
                depth = 0
                while True:
                    if s[j] == "<": depth += 1
                    elif s[j] == ">" and s[j - 1] != "-":
                        depth -= 1
                        if depth == 0: break
                    j += 1
                j += 1
            k = s.index("(", j)
            depth, e = 0, k
            while True:
                if s[e] == "(": depth += 1
                elif s[e] == ")":
                    depth -= 1
                    if depth == 0: break
                e += 1
            params_src = s[k + 1:e]
            b = e + 1
            # find body '{' (skip return type / where clause); a ';' first means a declaration without body
            while s[b] not in "{;":
                b += 1
            if s[b] == ";":
                continue
            depth, q = 0, b
            while True:
                if s[q] == "{": depth += 1
                elif s[q] == "}":
                    depth -= 1
                    if depth == 0: break
                q += 1
            self.fn_src[name] = (params_src, s[b:q + 1])

    def _use(self, spec):
        spec = spec.replace("\n", " ")
        m = re.match(r"(\w+)::\*$", spec)
        if m:
            self.globs.append(m.group(1)); return
        m = re.match(r"(\w+)::\{(.*)\}$", spec, re.S)
        if m:
            for item in m.group(2).split(","):
                item = item.strip()
                if item == "*": self.globs.append(m.group(1))
                elif item and "::" not in item and item != "self": self.explicit[item] = m.group(1)
            return
        m = re.match(r"(\w+)::(\w+)$", spec)
        if m:
            self.explicit[m.group(2)] = m.group(1); return
        m = re.match(r"\{(.*)\}$", spec, re.S)       # use crate::{native::{..}, utils::*}
        if m:
            depth, cur, parts = 0, "", []
            for ch in m.group(1):
                if ch == "{": depth += 1
                if ch == "}": depth -= 1
                if ch == "," and depth == 0:
                    parts.append(cur); cur = ""
                else:
                    cur += ch
            parts.append(cur)
            for part in parts:
                if part.strip(): self._use(part.strip())


def parse_params(src):
    names = []
    depth, cur, parts = 0, "", []
    for ch in src:
        if ch in "<([": depth += 1
        if ch in ">)]": depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur); cur = ""
        else:
            cur += ch
    parts.append(cur)
    for part in parts:
        part = part.strip()
        if not part: continue
        if part in ("&self", "self", "&mut self"):
            names.append("self"); continue
        nm = part.split(":")[0].strip()
        nm = re.sub(r"^(mut|ref)\s+", "", nm)
        names.append(nm)
    return names


class ReturnEx(Exception):
    def __init__(self, v): self.v = v


class Closure:
    def __init__(self, params, body, env, interp, file):
        self.params, self.body, self.env, self.interp, self.file = params, body, env, interp, file

    def __call__(self, *args):
        env = Env(self.env)
        for pat, a in zip(self.params, args):
            self.interp.bind(pat, a, env)
        return self.interp.ev(self.body, env, self.file)


class Env:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def get(self, k):
        e = self
        while e is not None:
            if k in e.vars: return e.vars[k]
            e = e.parent
        raise KeyError(k)

    def has(self, k):
        e = self
        while e is not None:
            if k in e.vars: return True
            e = e.parent
        return False

    def set_existing(self, k, v):
        e = self
        while e is not None:
            if k in e.vars:
                e.vars[k] = v; return
            e = e.parent
        raise KeyError(k)


class Some:
    __slots__ = ("v",)
    def __init__(self, v): self.v = v


class NStruct:
    """native.rs tuple structs Fp([u32;12]) / Fp2([Fp;2]) as far as the constraint code touches them."""
    def __init__(self, inner): self.inner = inner
    def get_u32_slice(self): return [x.inner if isinstance(x, NStruct) else x for x in self.inner]


class SelfObj:
    def __init__(self, **kw): self.__dict__.update(kw)


class Interp:
    def __init__(self, ref_src_dir):
        self.dir = ref_src_dir
        self.files = {}
        for fn in sorted(os.listdir(ref_src_dir)):
            if fn.endswith(".rs"):
                self.files[fn[:-3]] = RFile(fn[:-3], open(os.path.join(ref_src_dir, fn)).read())
        self.dag = Dag()
        self.constraints = []       # (class, node id)
        self.native_cache = {}
        self.call_depth = 0
        self.trace_calls = None     # optional callback(name, n_constraints_before, n_after, depth)

    # ---- constants ----
    def const(self, name, file):
        f = self.files[file]
        if name in f.const_val:
            return f.const_val[name]
        if name in f.const_src:
            ast = Parser(tokenize(f.const_src[name])).expr()
            v = self.ev(ast, Env(), file)
            f.const_val[name] = v
            return v
        if name in f.explicit and f.explicit[name] in self.files:
            return self.const(name, f.explicit[name])
        for g in f.globs:
            if g in self.files and (name in self.files[g].const_src):
                return self.const(name, g)
        raise KeyError("const %s not visible from %s.rs" % (name, file))

    def find_fn(self, name, file):
        f = self.files[file]
        if name in f.fn_src: return file
        if name in f.explicit and f.explicit[name] in self.files and name in self.files[f.explicit[name]].fn_src:
            return f.explicit[name]
        for g in f.globs:
            if g in self.files and name in self.files[g].fn_src:
                return g
        for g, rf in self.files.items():
            if name in rf.fn_src: return g
        return None

    def fn_ast(self, name, file):
        f = self.files[file]
        if name not in f.fn_ast:
            params_src, body_src = f.fn_src[name]
            body = Parser(tokenize(body_src)).block()
            f.fn_ast[name] = (parse_params(params_src), body)
        return f.fn_ast[name]

    # ---- native constants read from native.rs literals ----
    def native_strings(self, impl_fn):
        body = self.files["native"].fn_src[impl_fn][1]
        return [int(x) for x in re.findall(r'from_str\("(\d+)"\)', body)]

    def native_table(self, ty, fn):
        key = (ty, fn)
        if key in self.native_cache: return self.native_cache[key]
        src = self.files["native"].src
        # locate `impl <ty> {` block containing `fn <fn>`
        best = None
        for m in re.finditer(r"impl\s+%s\s*\{" % ty, src):
            j = src.find("fn %s(" % fn, m.end())
            if j >= 0:
                nxt = src.find("\nimpl ", m.end())
                if nxt < 0 or j < nxt:
                    best = j; break
        if best is None: raise KeyError("%s::%s" % (ty, fn))
        b = src.index("{", src.index(")", best))
        depth, q = 0, b
        while True:
            if src[q] == "{": depth += 1
            elif src[q] == "}":
                depth -= 1
                if depth == 0: break
            q += 1
        nums = [int(x) for x in re.findall(r'from_str\("(\d+)"\)', src[b:q])]
        limbs = lambda v: NStruct([(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)])
        if ty == "Fp2" and fn == "forbenius_coefficients":
            out = [limbs(v) for v in nums]
            assert len(out) == 2
        else:
            assert len(nums) % 2 == 0
            out = [NStruct([limbs(nums[2 * i]), limbs(nums[2 * i + 1])]) for i in range(len(nums) // 2)]
            assert len(out) in (6, 12)
        self.native_cache[key] = out
        return out

    # ---- evaluation ----
    def bind(self, pat, val, env):
        if pat[0] == "pname":
            env.vars[pat[1]] = val
        else:
            assert len(pat[1]) == len(val), (pat, val)
            for p, v in zip(pat[1], val):
                self.bind(p, v, env)

    def run_block(self, blk, env, file):
        env = Env(env)
        last = None
        for st in blk[1]:
            last = self.exec(st, env, file)
        return last

    def exec(self, st, env, file):
        k = st[0]
        if k == "let":
            v = self.ev(st[2], env, file) if st[2] is not None else None
            self.bind(st[1], v, env)
            return None
        if k == "expr":
            v = self.ev(st[1], env, file)
            return None if st[2] else v
        if k == "for":
            it = self.ev(st[2], env, file)
            for x in it:
                e2 = Env(env)
                self.bind(st[1], x, e2)
                self.run_block(st[3], e2, file)
            return None
        if k == "while":
            while self.ev(st[1], env, file):
                self.run_block(st[2], env, file)
            return None
        if k == "assign":
            op, lhs, rhs = st[1], st[2], self.ev(st[3], env, file)
            if lhs[0] == "path" and len(lhs[1]) == 1:
                name = lhs[1][0]
                if op != "=":
                    cur = env.get(name)
                    rhs = {"+=": cur + rhs, "-=": cur - rhs, "*=": cur * rhs}[op] if op in ("+=", "-=", "*=") else rhs
                env.set_existing(name, rhs)
                return None
            if lhs[0] == "index":
                base = self.ev(lhs[1], env, file)
                idx = self.ev(lhs[2], env, file)
                if op != "=":
                    rhs = {"+=": base[idx] + rhs, "-=": base[idx] - rhs, "*=": base[idx] * rhs}[op]
                base[idx] = rhs
                return None
            raise NotImplementedError("assign to %r" % (lhs,))
        if k == "return":
            raise ReturnEx(self.ev(st[1], env, file) if st[1] is not None else None)
        raise NotImplementedError(k)

    def call_fn(self, name, args, file):
        target = self.find_fn(name, file)
        if target is None:
            raise KeyError("fn %s (from %s.rs)" % (name, file))
        params, body = self.fn_ast(name, target)
        if len(params) != len(args):
            raise TypeError("%s expects %d args, got %d" % (name, len(params), len(args)))
        env = Env()
        for p, a in zip(params, args):
            env.vars[p] = a
        before = len(self.constraints)
        self.call_depth += 1
        try:
            ret = self.run_block(body, env, target)
        except ReturnEx as r:
            ret = r.v
        self.call_depth -= 1
        if self.trace_calls:
            self.trace_calls(name, before, len(self.constraints), self.call_depth)
        return ret

    def ev(self, e, env, file):
        k = e[0]
        if k == "lit":
            return e[1]
        if k == "path":
            return self.ev_path(e[1], env, file)
        if k == "bin":
            op = e[1]
            if op == "&&":
                return self.ev(e[2], env, file) and self.ev(e[3], env, file)
            if op == "||":
                return self.ev(e[2], env, file) or self.ev(e[3], env, file)
            a, b = self.ev(e[2], env, file), self.ev(e[3], env, file)
            if op == "+": return a + b
            if op == "-": return a - b
            if op == "*": return a * b
            if op == "/": return a // b
            if op == "%": return a % b
            if op == "<<": return a << b
            if op == ">>": return a >> b
            if op == "==": return a == b
            if op == "!=": return a != b
            if op == "<": return a < b
            if op == ">": return a > b
            if op == "<=": return a <= b
            if op == ">=": return a >= b
            if op == "&": return a & b
            if op == "|": return a | b
            if op == "^": return a ^ b
            raise NotImplementedError(op)
        if k == "un":
            v = self.ev(e[2], env, file)
            return (not v) if e[1] == "!" else -v
        if k == "cast":
            v = self.ev(e[1], env, file)
            if isinstance(v, bool): v = int(v)
            mask = {"u8": 0xFF, "u16": 0xFFFF, "u32": 0xFFFFFFFF, "u64": (1 << 64) - 1, "usize": (1 << 64) - 1}.get(e[2])
            return (v & mask) if (mask is not None and isinstance(v, int)) else v
        if k == "index":
            base = self.ev(e[1], env, file)
            idx = self.ev(e[2], env, file)
            if isinstance(idx, range) and isinstance(base, list):
                return base[idx.start:idx.stop]
            return base[idx]
        if k == "range":
            lo = self.ev(e[1], env, file) if e[1] is not None else 0
            hi = self.ev(e[2], env, file) if e[2] is not None else None
            if hi is None: return ("openrange", lo)
            return range(lo, hi + 1 if e[3] else hi)
        if k == "tuple":
            return tuple(self.ev(x, env, file) for x in e[1])
        if k == "array":
            return [self.ev(x, env, file) for x in e[1]]
        if k == "repeat":
            v = self.ev(e[1], env, file)
            return [v for _ in range(self.ev(e[2], env, file))]
        if k == "block":
            return self.run_block(e, env, file)
        if k == "if":
            if self.ev(e[1], env, file):
                return self.run_block(e[2], env, file)
            if e[3] is None: return None
            return self.ev(e[3], env, file) if e[3][0] == "if" else self.run_block(e[3], env, file)
        if k == "closure":
            return Closure(e[1], e[2], env, self, file)
        if k == "field":
            v = self.ev(e[1], env, file)
            if isinstance(e[2], int):
                if isinstance(v, NStruct): return v.inner
                return v[e[2]]
            return getattr(v, e[2])
        if k == "macro":
            return self.ev_macro(e[1], e[2], env, file)
        if k == "call":
            return self.ev_call(e[1], e[2], env, file)
        if k == "method":
            return self.ev_method(e[1], e[2], e[3], env, file)
        raise NotImplementedError(k)

    def ev_path(self, segs, env, file):
        if len(segs) == 1:
            n = segs[0]
            if env.has(n): return env.get(n)
            if n == "None": return None
            if n == "true": return True
            if n == "false": return False
            if n.isupper() or re.match(r"^[A-Z][A-Z0-9_]*$", n):
                return self.const(n, file)
            raise KeyError("unbound name %s in %s.rs" % (n, file))
        head, last = segs[0], segs[-1]
        if last in ("ONES", "ONE"): return Sym(self.dag, self.dag.const(1))
        if last in ("ZEROS", "ZERO"): return Sym(self.dag, self.dag.const(0))
        if last == "TWO": return Sym(self.dag, self.dag.const(2))
        if last == "NEG_ONE": return Sym(self.dag, self.dag.const(P - 1))
        if head in ("usize", "u32", "u64") and last == "MAX":
            return {"usize": (1 << 64) - 1, "u64": (1 << 64) - 1, "u32": (1 << 32) - 1}[head]
        if head == "crate" and len(segs) == 3 and segs[1] in self.files:
            return self.const(last, segs[1])
        if head in self.files and len(segs) == 2:
            return self.const(last, head)
        return ("fnpath", tuple(segs))

    def ev_macro(self, name, items, env, file):
        if name == "bit_decomp_32":
            row = self.ev(items[0], env, file)
            col = self.ev(items[1], env, file)
            acc = Sym(self.dag, self.dag.const(0))
            for i in range(32):          # (0..32).fold(P::ZEROS, |acc, i| acc + row[col + i] * FE::from_canonical_u64(1 << i))
                acc = acc + row[col + i] * Sym(self.dag, self.dag.const(1 << i))
            return acc
        if name == "vec":
            if len(items) == 1 and items[0][0] == "repeat":
                return self.ev(items[0], env, file)
            return [self.ev(x, env, file) for x in items]
        if name in ("assert", "assert_eq", "debug_assert", "println", "debug_assert_eq"):
            return None
        raise NotImplementedError("macro %s!" % name)

    def ev_call(self, fn, args_ast, env, file):
        args = [self.ev(a, env, file) for a in args_ast]
        if fn[0] == "path":
            segs = fn[1]
            if len(segs) == 1:
                n = segs[0]
                if env.has(n) and callable(env.get(n)):
                    return env.get(n)(*args)
                if n == "Some": return Some(args[0])
                if n == "Ok": return args[0]
                return self.builtin_or_user(n, args, file)
            head, last = segs[0], segs[-1]
            if last.startswith("from_canonical_") or last in ("from_noncanonical_u64",):
                v = args[0]
                if isinstance(v, Sym): return v
                return Sym(self.dag, self.dag.const(int(v)))
            if head == "BigUint":
                if last in ("from", "new"):
                    v = args[0]
                    if isinstance(v, list): return sum(int(x) << (32 * i) for i, x in enumerate(v))
                    return int(v)
                if last == "from_str": return int(args[0])
            if head in ("Fp2", "Fp6", "Fp12") and last.startswith("forbenius_coefficients"):
                return self.native_table(head, last)
            if head in ("Vec",) and last == "new": return []
            if len(segs) == 2 and head in self.files:
                return self.call_fn(last, args, head)
            if head == "crate" and len(segs) == 3:
                return self.call_fn(last, args, segs[1])
            if head == "Self" or head == "self":
                return self.call_fn(last, args, file)
            raise NotImplementedError("call %s" % "::".join(segs))
        f = self.ev(fn, env, file)
        return f(*args)

    def builtin_or_user(self, n, args, file):
        if n == "modulus":
            return int(re.search(r'from_str\("(\d+)"\)', self.files["native"].fn_src["modulus"][1]).group(1))
        if n in ("get_u32_vec_from_literal", "get_u32_vec_from_literal_ref"):
            return [(args[0] >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
        if n in ("get_u32_vec_from_literal_24", "get_u32_vec_from_literal_ref_24"):
            return [(args[0] >> (32 * i)) & 0xFFFFFFFF for i in range(24)]
        if n == "mod_inverse":
            return pow(args[0], -1, args[1])
        if n == "get_bls_12_381_parameter":
            return int(re.search(r'from_str\("(\d+)"\)', self.files["native"].fn_src[n][1]).group(1))
        if n == "log2_ceil":
            return (args[0] - 1).bit_length()
        return self.call_fn(n, args, file)

    def ev_method(self, recv_ast, name, args_ast, env, file):
        # constraint emission
        if name in ("constraint", "constraint_transition", "constraint_first_row", "constraint_last_row") \
                and recv_ast[0] == "path" and recv_ast[1][-1] in ("yield_constr", "consumer"):
            v = self.ev(args_ast[0], env, file)
            if isinstance(v, int): v = Sym(self.dag, self.dag.const(v))
            cls = {"constraint": 1, "constraint_transition": 2, "constraint_first_row": 3, "constraint_last_row": 4}[name]
            self.constraints.append((cls, v.id))
            return None
        recv = self.ev(recv_ast, env, file)
        args = [self.ev(a, env, file) for a in args_ast]
        if name == "unwrap_or":
            return recv.v if isinstance(recv, Some) else args[0]
        if name in ("unwrap", "expect"):
            return recv.v if isinstance(recv, Some) else recv
        if name in ("iter", "into_iter", "clone", "to_vec", "try_into", "into", "copied", "cloned", "to_owned", "as_slice"):
            return recv
        if name == "is_some": return isinstance(recv, Some)
        if name == "is_none": return recv is None
        if name == "map":
            if isinstance(recv, Some): return Some(args[0](recv.v))
            if recv is None: return None
            return [args[0](x) for x in recv]
        if name == "collect":
            return list(recv)
        if name == "rev": return list(recv)[::-1]
        if name == "enumerate": return [(i, x) for i, x in enumerate(recv)]
        if name == "zip": return [(a, b) for a, b in zip(recv, args[0])]
        if name == "step_by": return list(recv)[::args[0]]
        if name == "fold":
            acc = args[0]
            for x in recv: acc = args[1](acc, x)
            return acc
        if name == "sum":
            items = list(recv)
            acc = items[0]
            for x in items[1:]: acc = acc + x
            return acc
        if name == "len": return len(recv)
        if name == "concat": return [y for x in recv for y in x]
        if name == "pow": return recv ** args[0]
        if name == "to_u32_digits":
            out, v = [], recv
            while v: out.append(v & 0xFFFFFFFF); v >>= 32
            return out
        if name == "get_u32_slice": return recv.get_u32_slice()
        if name == "get_local_values": return recv.local
        if name == "get_next_values": return recv.next
        if name == "get_public_inputs": return recv.pis
        if name == "push":
            recv.append(args[0]); return None
        if name == "exp_u64": raise NotImplementedError("exp on symbolic values")
        raise NotImplementedError("method .%s on %r" % (name, type(recv)))

    # ---- entry point ----
    def trace_stark(self, file, n_cols, n_pis, **self_fields):
        vars_obj = SelfObj(local=ColVec(self.dag, LOCAL, n_cols), next=ColVec(self.dag, NEXT, n_cols),
                           pis=ColVec(self.dag, PI, n_pis))
        params, body = self.fn_ast("eval_packed_generic", file)
        env = Env()
        assert params == ["self", "vars", "yield_constr"], params
        env.vars["self"] = SelfObj(**self_fields)
        env.vars["vars"] = vars_obj
        env.vars["yield_constr"] = None
        try:
            self.run_block(body, env, file)
        except ReturnEx:
            pass
        return self.constraints
