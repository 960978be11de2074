"""A Yul-subset interpreter with the four BN254/modexp precompiles, enough to execute the reference's
generated verifier (proving-server/P256Verifier.yul, object "Runtime") on a proof.
TEST INFRASTRUCTURE ONLY.

Supported: function definitions, blocks, `let` (typed or not), `:=`, `if`, literals, and the builtins
the file uses — mstore / mstore8 / mload / calldataload / keccak256 / mulmod / addmod / mod / add / sub /
lt / eq / and / not / gas / staticcall (precompiles 0x5 modexp, 0x6 ecAdd, 0x7 ecMul, 0x8 ecPairing) /
revert / return.  `not` on the typed-bool `success` flag is the logical negation (the file is written
in snark-verifier's typed Yul: `let success:bool := true … if not(success) { revert(0, 0) }`).
"""
from __future__ import annotations

import re

from . import pairing
from .keccak import keccak256
from .pyref import P, g1_add, g1_is_on_curve, g1_mul

M256 = (1 << 256) - 1


class Revert(Exception):
    pass


class Return(Exception):
    def __init__(self, data: bytes):
        self.data = data


_TOKEN = re.compile(r"\s*(?:(//[^\n]*)|(0x[0-9a-fA-F]+|\d+)|([A-Za-z_][A-Za-z_0-9]*)|(:=|[{}(),:])|(\"[^\"]*\"))")


def tokenize(src: str):
    pos, out = 0, []
    while True:
        m = _TOKEN.match(src, pos)
        if not m:
            if src[pos:].strip():
                raise SyntaxError(f"yul: cannot tokenize at {pos}: {src[pos:pos+40]!r}")
            return out
        pos = m.end()
        if m.group(1):
            continue
        if m.group(2):
            out.append(("num", int(m.group(2), 0)))
        elif m.group(3):
            out.append(("id", m.group(3)))
        elif m.group(4):
            out.append(("sym", m.group(4)))
        else:
            out.append(("str", m.group(5)[1:-1]))


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, kind, val=None):
        tok = self.next()
        if tok[0] != kind or (val is not None and tok[1] != val):
            raise SyntaxError(f"yul: expected {kind} {val}, got {tok}")
        return tok

    def skip_type(self):
        if self.peek() == ("sym", ":"):
            self.next()
            self.expect("id")

    def block(self):
        self.expect("sym", "{")
        stmts = []
        while self.peek() != ("sym", "}"):
            stmts.append(self.statement())
        self.expect("sym", "}")
        return ("block", stmts)

    def statement(self):
        k, v = self.peek()
        if (k, v) == ("sym", "{"):
            return self.block()
        if k == "id" and v == "function":
            self.next()
            name = self.expect("id")[1]
            self.expect("sym", "(")
            params = []
            while self.peek() != ("sym", ")"):
                params.append(self.expect("id")[1])
                self.skip_type()
                if self.peek() == ("sym", ","):
                    self.next()
            self.expect("sym", ")")
            rets = []
            if self.peek() == ("sym", "-"):
                pass
            # "-> r:type"
            if self.peek()[0] == "id" and False:
                pass
            if self._arrow():
                while True:
                    rets.append(self.expect("id")[1])
                    self.skip_type()
                    if self.peek() == ("sym", ","):
                        self.next()
                        continue
                    break
            body = self.block()
            return ("func", name, params, rets, body)
        if k == "id" and v == "let":
            self.next()
            names = [self.expect("id")[1]]
            self.skip_type()
            while self.peek() == ("sym", ","):
                self.next()
                names.append(self.expect("id")[1])
                self.skip_type()
            val = None
            if self.peek() == ("sym", ":="):
                self.next()
                val = self.expr()
            return ("let", names, val)
        if k == "id" and v == "if":
            self.next()
            cond = self.expr()
            return ("if", cond, self.block())
        if k == "id":
            # assignment or expression statement
            if self.t[self.i + 1] == ("sym", ":="):
                name = self.next()[1]
                self.next()
                return ("assign", name, self.expr())
            return ("expr", self.expr())
        raise SyntaxError(f"yul: unexpected token {self.peek()}")

    def _arrow(self):
        # the tokenizer drops "->" (not in its symbol set) — detect it from the raw token stream marker
        if self.peek() == ("arrow", None):
            self.next()
            return True
        return False

    def expr(self):
        k, v = self.next()
        if k == "num":
            return ("num", v)
        if k == "id":
            if v == "true":
                return ("num", 1)
            if v == "false":
                return ("num", 0)
            if self.peek() == ("sym", "("):
                self.next()
                args = []
                while self.peek() != ("sym", ")"):
                    args.append(self.expr())
                    if self.peek() == ("sym", ","):
                        self.next()
                self.expect("sym", ")")
                return ("call", v, args)
            self.skip_type()
            return ("var", v)
        raise SyntaxError(f"yul: bad expression token {(k, v)}")


def parse_runtime(yul_source: str):
    """Extract object "Runtime" { code { ... } } and parse its code block."""
    idx = yul_source.index('object "Runtime"')
    start = yul_source.index("code", idx)
    brace = yul_source.index("{", start)
    depth, j = 0, brace
    while True:
        ch = yul_source[j]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    body = yul_source[brace:j + 1].replace("->", " __arrow__ ")
    toks = [("arrow", None) if t == ("id", "__arrow__") else t for t in tokenize(body)]
    return Parser(toks).block()


class Machine:
    def __init__(self, calldata: bytes):
        self.calldata = calldata
        self.mem = bytearray()
        self.funcs = {}
        self.precompile_calls = {5: 0, 6: 0, 7: 0, 8: 0}
        self.keccak_calls = 0

    # ---- memory ----
    def _grow(self, end):
        if len(self.mem) < end:
            self.mem.extend(b"\0" * (end - len(self.mem)))

    def mload(self, p):
        self._grow(p + 32)
        return int.from_bytes(self.mem[p:p + 32], "big")

    def mstore(self, p, v):
        self._grow(p + 32)
        self.mem[p:p + 32] = (v & M256).to_bytes(32, "big")

    def mread(self, p, n):
        self._grow(p + n)
        return bytes(self.mem[p:p + n])

    # ---- precompiles ----
    def staticcall(self, addr, inp, insz, outp, outsz):
        data = self.mread(inp, insz)
        self.precompile_calls[addr] = self.precompile_calls.get(addr, 0) + 1
        word = lambda i: int.from_bytes(data[32 * i:32 * i + 32], "big")
        try:
            if addr == 5:  # modexp: lengths then base, exp, mod
                bl, el, ml = word(0), word(1), word(2)
                base = int.from_bytes(data[96:96 + bl], "big")
                exp = int.from_bytes(data[96 + bl:96 + bl + el], "big")
                mod = int.from_bytes(data[96 + bl + el:96 + bl + el + ml], "big")
                res = (pow(base, exp, mod) if mod else 0).to_bytes(ml, "big")
            elif addr == 6:  # ecAdd
                a, b = self._g1(word(0), word(1)), self._g1(word(2), word(3))
                res = self._enc(g1_add(a, b))
            elif addr == 7:  # ecMul
                a = self._g1(word(0), word(1))
                res = self._enc(g1_mul_full(a, word(2)))
            elif addr == 8:  # ecPairing
                if insz % 192:
                    return 0
                pairs = []
                for k in range(insz // 192):
                    w = [word(6 * k + i) for i in range(6)]
                    g1 = self._g1(w[0], w[1])
                    x = (w[3], w[2])  # EIP-197: imaginary part first
                    y = (w[5], w[4])
                    g2 = None if (x == (0, 0) and y == (0, 0)) else (x, y)
                    if g2 is not None and not pairing.g2_is_on_curve(g2):
                        return 0
                    pairs.append((g1, g2))
                res = (1 if pairing.pairing_product_is_one(pairs) else 0).to_bytes(32, "big")
            else:
                return 0
        except ValueError:
            return 0
        res = res[:outsz].ljust(outsz, b"\0") if len(res) < outsz else res[:outsz]
        self._grow(outp + outsz)
        self.mem[outp:outp + outsz] = res
        return 1

    @staticmethod
    def _g1(x, y):
        if x == 0 and y == 0:
            return None
        if x >= P or y >= P or not g1_is_on_curve((x, y)):
            raise ValueError("invalid G1 point")
        return (x, y)

    @staticmethod
    def _enc(pt):
        if pt is None:
            return b"\0" * 64
        return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")

    # ---- evaluation ----
    def builtin(self, name, a):
        if name == "mload":
            return self.mload(a[0])
        if name == "mstore":
            self.mstore(a[0], a[1]); return None
        if name == "mstore8":
            self._grow(a[0] + 1); self.mem[a[0]] = a[1] & 0xFF; return None
        if name == "calldataload":
            chunk = self.calldata[a[0]:a[0] + 32]
            return int.from_bytes(chunk.ljust(32, b"\0"), "big")
        if name == "keccak256":
            self.keccak_calls += 1
            return int.from_bytes(keccak256(self.mread(a[0], a[1])), "big")
        if name == "mulmod":
            return (a[0] * a[1]) % a[2] if a[2] else 0
        if name == "addmod":
            return (a[0] + a[1]) % a[2] if a[2] else 0
        if name == "mod":
            return a[0] % a[1] if a[1] else 0
        if name == "add":
            return (a[0] + a[1]) & M256
        if name == "sub":
            return (a[0] - a[1]) & M256
        if name == "lt":
            return int(a[0] < a[1])
        if name == "eq":
            return int(a[0] == a[1])
        if name == "and":
            return a[0] & a[1]
        if name == "not":
            return int(not a[0]) if a[0] in (0, 1) else (~a[0]) & M256  # typed-bool `not(success)`
        if name == "gas":
            return M256
        if name == "staticcall":
            return self.staticcall(a[1], a[2], a[3], a[4], a[5])
        if name == "revert":
            raise Revert()
        if name == "return":
            raise Return(self.mread(a[0], a[1]))
        raise NameError(f"yul: unknown builtin {name}")

    def eval(self, e, scopes):
        kind = e[0]
        if kind == "num":
            return e[1]
        if kind == "var":
            for s in reversed(scopes):
                if e[1] in s:
                    return s[e[1]]
            raise NameError(e[1])
        name, args = e[1], [self.eval(x, scopes) for x in e[2]]
        if name in self.funcs:
            _, _, params, rets, body = self.funcs[name]
            frame = dict(zip(params, args))
            for r in rets:
                frame[r] = 0
            self.exec_block(body, [frame])
            return frame[rets[0]] if rets else None
        return self.builtin(name, args)

    def exec_block(self, blk, scopes):
        scopes = scopes + [{}]
        for st in blk[1]:
            if st[0] == "func":
                self.funcs[st[1]] = st
        for st in blk[1]:
            k = st[0]
            if k == "func":
                continue
            if k == "block":
                self.exec_block(st, scopes)
            elif k == "let":
                val = self.eval(st[2], scopes) if st[2] is not None else 0
                scopes[-1][st[1][0]] = val
            elif k == "assign":
                val = self.eval(st[2], scopes)
                for s in reversed(scopes):
                    if st[1] in s:
                        s[st[1]] = val
                        break
                else:
                    raise NameError(st[1])
            elif k == "if":
                if self.eval(st[1], scopes):
                    self.exec_block(st[2], scopes)
            elif k == "expr":
                self.eval(st[1], scopes)


def g1_mul_full(pt, k: int):
    """ecMul takes any 256-bit scalar (no reduction needed for correctness: group order r)."""
    from .pyref import R
    return g1_mul(pt, k % R)


def run_verifier(yul_source: str, calldata: bytes):
    """Returns (accepted: bool, machine).  accepted == the call did not revert."""
    ast = parse_runtime(yul_source)
    m = Machine(calldata)
    try:
        m.exec_block(ast, [])
    except Return:
        return True, m
    except Revert:
        return False, m
    return True, m
