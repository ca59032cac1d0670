"""A small tree-walking interpreter for the Lua subset the glue in lua/radio_b200/ is written in, on top of the lark parse
tree of tests/lua_grammar.py.  LuaJIT is not installed in the build image; with this the glue's LOGIC -- the scheduler's
`collapse_gpu_runs`, the `b200.install` class patching, the per-block create calls -- is EXECUTED in the CPU test-suite
against mock `radio` / `ffi` / library objects instead of only being read.

Covered: locals and upvalues (closures), assignment to names / fields / indices, multiple assignment and multiple returns,
`...`, if / while / repeat / numeric and generic `for`, `break`, `return`, function and method definitions, method calls,
table constructors, metatables (`__index` table or function, `__newindex` function, `__call`, the arithmetic events), arithmetic / comparison / logic / `..` / `#`,
and the handful of standard functions the glue uses.  `goto` to a label of an enclosing block.  Not covered (not used by the glue or the reference core): coroutines,
string methods via `:`, integer division, metamethods other than `__index` / `__call` / `__newindex`-less tables."""
from lark import Token, Tree

from tests.lua_grammar import parse_lua


class LuaError(Exception):
    pass


class LuaTable:
    __slots__ = ("hash", "meta")

    def __init__(self, items=None):
        self.hash = dict(items or {})
        self.meta = None

    def __repr__(self):
        return "LuaTable(%d)" % len(self.hash)

    def length(self):
        n = 0
        while (n + 1) in self.hash:
            n += 1
        return n

    def array(self):
        return [self.hash[i] for i in range(1, self.length() + 1)]


def _key(k):
    if isinstance(k, float) and k.is_integer():
        return int(k)
    return k


class _Break(Exception):
    pass


class _Goto(Exception):
    def __init__(self, label):
        self.label = label


class _Return(Exception):
    def __init__(self, values):
        self.values = values


class Scope:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def lookup(self, name):
        s = self
        while s is not None:
            if name in s.vars:
                return s
            s = s.parent
        return None


class LuaFunction:
    def __init__(self, interp, funcbody, scope, method, name="?"):
        self.interp, self.scope, self.name = interp, scope, name
        kids = funcbody.children
        self.block = kids[-1]
        self.params, self.vararg = (["self"] if method else []), False
        for c in kids[:-1]:
            if isinstance(c, Tree) and c.data == "parlist":
                for cc in c.children:
                    if isinstance(cc, Tree) and cc.data == "namelist":
                        self.params += [str(t) for t in cc.children]
                    elif isinstance(cc, Token) and cc.type == "VARARG":
                        self.vararg = True

    def __call__(self, *args):
        sc = Scope(self.scope)
        for i, p in enumerate(self.params):
            sc.vars[p] = args[i] if i < len(args) else None
        sc.vars["..."] = list(args[len(self.params):]) if self.vararg else []
        try:
            self.interp.exec_block(self.block, sc)
        except _Return as r:
            return r.values
        return []


class Interp:
    def __init__(self, modules=None, globals_=None):
        self.modules = dict(modules or {})        # name -> value, or name -> source text (str) to be run on demand
        self.loaded = {}
        self.G = Scope()
        self.G.vars.update(self.stdlib())
        self.G.vars.update(globals_ or {})

    # ---- standard library subset -------------------------------------------------------------------------------------
    def stdlib(self):
        def lua_pairs(t):
            return [lambda tt, k: None, t, None]      # marker; generic-for special-cases pairs/ipairs results

        def lua_error(msg=None, *a):
            raise LuaError(self.tostring(msg))

        def lua_assert(v=None, msg="assertion failed!", *rest):
            if v is None or v is False:
                raise LuaError(self.tostring(msg))
            return [v, msg] + list(rest)

        def setmt(t, mt):
            t.meta = mt
            return [t]

        def tonumber(v=None, *a):
            if isinstance(v, (int, float)) and not isinstance(v, bool):
                return [v]
            try:
                f = float(v)
                return [int(f) if f.is_integer() else f]
            except (TypeError, ValueError):
                return [None]

        def lua_type(v=None):
            if v is None:
                return ["nil"]
            if isinstance(v, bool):
                return ["boolean"]
            if isinstance(v, (int, float)):
                return ["number"]
            if isinstance(v, str):
                return ["string"]
            if isinstance(v, LuaTable):
                return ["table"]
            if callable(v):
                return ["function"]
            return ["userdata"]

        def tinsert(t, *a):
            if len(a) == 1:
                t.hash[t.length() + 1] = a[0]
            else:
                pos, v = int(a[0]), a[1]
                n = t.length()
                for i in range(n, pos - 1, -1):
                    t.hash[i + 1] = t.hash[i]
                t.hash[pos] = v
            return []

        def tremove(t, pos=None):
            n = t.length()
            if n == 0:
                return [None]
            pos = n if pos is None else int(pos)
            v = t.hash.get(pos)
            for i in range(pos, n):
                t.hash[i] = t.hash[i + 1]
            t.hash.pop(n, None)
            return [v]

        def unpack(t, i=1, j=None):
            j = t.length() if j is None else int(j)
            return [t.hash.get(k) for k in range(int(i), j + 1)]

        def select(n, *a):
            if n == "#":
                return [len(a)]
            return list(a[int(n) - 1:])

        def fmt(f, *a):
            return [f % tuple(a)]

        import os as _os
        import math as _math
        import re as _re

        def lua_format(f, *a):
            # Lua's %s applies tostring(); %d wants an integer-valued number
            args, i = list(a), 0
            out = []
            for m in _re.finditer(r"%(?:%|[-+ #0]*\d*(?:\.\d+)?[a-zA-Z])|[^%]+", f):
                tok = m.group(0)
                if not tok.startswith("%") or tok == "%%":
                    out.append("%" if tok == "%%" else tok)
                    continue
                v = args[i] if i < len(args) else None
                i += 1
                conv = tok[-1]
                if conv == "s":
                    out.append(tok % self.tostring(v))
                elif conv in "di":
                    out.append((tok[:-1] + "d") % int(v))
                elif conv == "q":
                    out.append('"%s"' % v)
                else:
                    out.append(tok % v)
            return ["".join(out)]

        def lua_gsub(s, pat, repl):
            # plain-text patterns only (what the hot-path files use: "\n" -> ...)
            if any(c in pat for c in "^$()%.[]*+-?"):
                raise LuaError("string.gsub: pattern '%s' not supported by the test interpreter" % pat)
            return [s.replace(pat, repl), s.count(pat)]

        def lua_pcall(f, *a):
            try:
                return [True] + self.call(f, list(a))
            except LuaError as e:
                return [False, e.args[0]]

        def lua_next(t, k=None):
            keys = list(t.hash)
            if k is None:
                return [keys[0], t.hash[keys[0]]] if keys else [None]
            i = keys.index(_key(k)) + 1
            return [keys[i], t.hash[keys[i]]] if i < len(keys) else [None]

        def rawset(t, k, v):
            t.hash[_key(k)] = v
            return [t]

        table = LuaTable({"insert": tinsert, "remove": tremove, "unpack": unpack,
                          "concat": lambda t, sep="", *a: [sep.join(self.tostring(x) for x in t.array())]})
        string = LuaTable({"format": lua_format, "len": lambda s: [len(s)], "sub": lambda s, i, j=-1: [s[int(i) - 1:(None if j == -1 else int(j))]],
                           "gsub": lua_gsub, "rep": lambda s, n, *a: [s * int(n)], "lower": lambda s: [s.lower()], "upper": lambda s: [s.upper()],
                           "byte": lambda s, i=1: [ord(s[int(i) - 1])], "char": lambda *a: ["".join(chr(int(c)) for c in a)]})
        os_t = LuaTable({"getenv": lambda k: [_os.environ.get(k)], "exit": lambda *a: lua_error("os.exit called"), "time": lambda *a: [0]})
        mf = lambda f: (lambda *a: [f(*a)])
        math_t = LuaTable({"floor": lambda x: [_math.floor(x)], "ceil": lambda x: [_math.ceil(x)], "min": lambda *a: [min(a)], "max": lambda *a: [max(a)],
                           "huge": float("inf"), "pi": _math.pi, "sin": mf(_math.sin), "cos": mf(_math.cos), "tan": mf(_math.tan), "atan2": mf(_math.atan2),
                           "atan": mf(_math.atan), "sqrt": mf(_math.sqrt), "abs": mf(abs), "exp": mf(_math.exp), "log": mf(_math.log),
                           "log10": mf(_math.log10), "pow": mf(pow), "fmod": mf(_math.fmod), "sinh": mf(_math.sinh), "cosh": mf(_math.cosh)})
        stream = lambda: LuaTable({"write": lambda self_, *a: [self_], "flush": lambda self_: []})
        io_t = LuaTable({"stderr": stream(), "stdout": stream()})
        return {"pairs": lambda t: ["__pairs__", t], "ipairs": lambda t: ["__ipairs__", t], "error": lua_error, "assert": lua_assert,
                "setmetatable": setmt, "getmetatable": lambda t: [t.meta if isinstance(t, LuaTable) else None], "tonumber": tonumber,
                "tostring": lambda v=None: [self.tostring(v)], "type": lua_type, "unpack": unpack, "select": select, "print": lambda *a: [],
                "table": table, "string": string, "os": os_t, "math": math_t, "io": io_t, "require": lambda name: [self.require(name)],
                "rawget": lambda t, k: [t.hash.get(_key(k))], "rawset": rawset, "rawequal": lambda a, b: [a is b or a == b], "next": lua_next,
                "pcall": lua_pcall}

    def tostring(self, v):
        if v is None:
            return "nil"
        if v is True:
            return "true"
        if v is False:
            return "false"
        if isinstance(v, float) and v.is_integer():
            return str(int(v))
        return str(v)

    def require(self, name):
        if name in self.loaded:
            return self.loaded[name]
        if name not in self.modules:
            raise LuaError("module '%s' not found" % name)
        m = self.modules[name]
        if isinstance(m, str):
            vals = self.run(m, name)
            m = vals[0] if vals else True
        self.loaded[name] = m
        return m

    # ---- running -----------------------------------------------------------------------------------------------------
    def run(self, text, chunkname="chunk"):
        tree = parse_lua(text)
        block = tree.children[0]
        sc = Scope(self.G)
        sc.vars["..."] = []
        try:
            self.exec_block(block, sc)
        except _Return as r:
            return r.values
        return []

    def call(self, f, args):
        if isinstance(f, LuaTable):
            mt = f.meta
            h = mt.hash.get("__call") if mt else None
            if h is None:
                raise LuaError("attempt to call a table value")
            return self.call(h, [f] + list(args))
        if f is None:
            raise LuaError("attempt to call a nil value")
        if not callable(f):
            raise LuaError("attempt to call a %s value" % type(f).__name__)
        r = f(*args)
        if r is None:
            return []
        return r if isinstance(r, list) else [r]

    def index(self, obj, key):
        key = _key(key)
        if isinstance(obj, LuaTable):
            t = obj
            for _ in range(100):
                if key in t.hash:
                    return t.hash[key]
                mt = t.meta
                h = mt.hash.get("__index") if mt else None
                if h is None:
                    return None
                if isinstance(h, LuaTable):
                    t = h
                    continue
                return (self.call(h, [t, key]) or [None])[0]
            raise LuaError("__index chain too long")
        if isinstance(obj, str):
            return self.index(self.G.vars["string"], key)
        if obj is None:
            raise LuaError("attempt to index a nil value (field '%s')" % (key,))
        # a Python mock object: attribute access
        v = getattr(obj, str(key), None)
        return v

    def setindex(self, obj, key, value):
        if isinstance(obj, LuaTable):
            key = _key(key)
            if key not in obj.hash and obj.meta is not None and obj.meta.hash.get("__newindex") is not None:
                self.call(obj.meta.hash["__newindex"], [obj, key, value])
                return
            if value is None:
                obj.hash.pop(key, None)
            else:
                obj.hash[key] = value
        elif obj is None:
            raise LuaError("attempt to index a nil value (assign field '%s')" % (key,))
        else:
            setattr(obj, str(key), value)

    # ---- statements --------------------------------------------------------------------------------------------------
    def exec_block(self, block, scope):
        sc = Scope(scope)
        stats = block.children
        i = 0
        while i < len(stats):
            st = stats[i]
            i += 1
            try:
                self.exec_stat(st, sc)
            except _Goto as g:
                # `goto label` (LuaJIT): continue after the label if it is a statement of THIS block, else keep unwinding
                # (the reference only jumps forward to a `::continue::` at the end of a loop body, composite.lua:194-210)
                at = [k for k, s2 in enumerate(stats) if isinstance(s2, Tree) and s2.data == "label" and str(s2.children[0]) == g.label]
                if not at:
                    raise
                i = at[0] + 1
            except LuaError as e:
                if not getattr(e, "located", False) and isinstance(st, Tree) and getattr(st.meta, "line", None):
                    e.args = ("%s (line %d)" % (e.args[0], st.meta.line),)
                    e.located = True
                raise

    def exec_stat(self, st, sc):
        if isinstance(st, Token):
            return
        d = st.data
        if d == "local_assign":
            names = [str(t) for t in st.children[0].children]
            vals = self.eval_list(st.children[1], sc) if len(st.children) > 1 else []
            for i, n in enumerate(names):
                sc.vars[n] = vals[i] if i < len(vals) else None
        elif d == "assign":
            targets = st.children[0].children
            vals = self.eval_list(st.children[1], sc)
            for i, t in enumerate(targets):
                self.assign(t, vals[i] if i < len(vals) else None, sc)
        elif d == "functioncall":
            self.eval_call(st, sc)
        elif d == "local_function":
            name = str(st.children[0])
            sc.vars[name] = None
            sc.vars[name] = LuaFunction(self, st.children[1], sc, False, name)
        elif d == "function_stat":
            fn = st.children[0]
            toks = [str(t) for t in fn.children if isinstance(t, Token)]
            method = [c for c in fn.children if isinstance(c, Tree) and c.data == "method"]
            f = LuaFunction(self, st.children[1], sc, bool(method), ".".join(toks))
            path = toks + ([str(method[0].children[0])] if method else [])
            if len(path) == 1:
                s = sc.lookup(path[0]) or self.G
                s.vars[path[0]] = f
            else:
                s = sc.lookup(path[0])
                obj = s.vars[path[0]] if s else None
                for k in path[1:-1]:
                    obj = self.index(obj, k)
                self.setindex(obj, path[-1], f)
        elif d == "if_stat":
            kids = st.children
            i = 0
            while i + 1 < len(kids):
                if self.truthy(self.eval(kids[i], sc)):
                    self.exec_block(kids[i + 1], sc)
                    return
                i += 2
            if i < len(kids):
                self.exec_block(kids[i], sc)
        elif d == "while_loop":
            try:
                while self.truthy(self.eval(st.children[0], sc)):
                    self.exec_block(st.children[1], sc)
            except _Break:
                pass
        elif d == "repeat_loop":
            try:
                while True:
                    inner = Scope(sc)
                    for s2 in st.children[0].children:
                        self.exec_stat(s2, inner)
                    if self.truthy(self.eval(st.children[1], inner)):
                        break
            except _Break:
                pass
        elif d == "for_num":
            name = str(st.children[0])
            exps = [self.eval(e, sc) for e in st.children[1:-1]]
            start, stop, step = exps[0], exps[1], (exps[2] if len(exps) > 2 else 1)
            i = start
            try:
                while (step > 0 and i <= stop) or (step < 0 and i >= stop):
                    inner = Scope(sc)
                    inner.vars[name] = i
                    self.exec_block(st.children[-1], inner)
                    i += step
            except _Break:
                pass
        elif d == "for_in":
            names = [str(t) for t in st.children[0].children]
            vals = self.eval_list(st.children[1], sc)
            try:
                if vals and vals[0] in ("__pairs__", "__ipairs__"):
                    t = vals[1]
                    items = [(i, t.hash[i]) for i in range(1, t.length() + 1)] if vals[0] == "__ipairs__" else list(t.hash.items())
                    keys0 = set(t.hash)
                    for k, v in items:
                        if vals[0] == "__pairs__" and k not in t.hash:
                            continue                    # cleared during the traversal (allowed)
                        inner = Scope(sc)
                        for n, x in zip(names, (k, v)):
                            inner.vars[n] = x
                        for n in names[2:]:
                            inner.vars[n] = None
                        self.exec_block(st.children[2], inner)
                        # "The behavior of next is undefined if, during the traversal, you assign any value to a
                        # non-existent field" (Lua 5.1 manual, next): here it is an error, so the tests catch it
                        if vals[0] == "__pairs__" and not set(t.hash) <= keys0:
                            raise LuaError("new key assigned to a table during its pairs() traversal")
                else:
                    f, s, ctl = (vals + [None, None, None])[:3]
                    while True:
                        r = self.call(f, [s, ctl])
                        if not r or r[0] is None:
                            break
                        ctl = r[0]
                        inner = Scope(sc)
                        for i, n in enumerate(names):
                            inner.vars[n] = r[i] if i < len(r) else None
                        self.exec_block(st.children[2], inner)
            except _Break:
                pass
        elif d == "do_block":
            self.exec_block(st.children[0], sc)
        elif d == "return_stat":
            raise _Return(self.eval_list(st.children[0], sc) if st.children else [])
        elif d == "break_stat":
            raise _Break()
        elif d == "stat" or d == "label":
            return
        elif d == "goto_stat":
            raise _Goto(str(st.children[0]))
        else:
            raise LuaError("statement not supported by the test interpreter: %s" % d)

    def assign(self, target, value, sc):
        kids = target.children
        if len(kids) == 1:
            name = str(kids[0])
            s = sc.lookup(name) or self.G
            s.vars[name] = value
        else:
            obj = self.eval(kids[0], sc)
            key = str(kids[1]) if isinstance(kids[1], Token) and kids[1].type == "NAME" else self.eval(kids[1], sc)
            self.setindex(obj, key, value)

    # ---- expressions -------------------------------------------------------------------------------------------------
    @staticmethod
    def truthy(v):
        return v is not None and v is not False

    def eval_list(self, explist, sc):
        out = []
        kids = explist.children if isinstance(explist, Tree) and explist.data == "explist" else [explist]
        for i, e in enumerate(kids):
            if i == len(kids) - 1:
                out.extend(self.eval_multi(e, sc))
            else:
                out.append(self.eval(e, sc))
        return out

    def eval_multi(self, e, sc):
        if isinstance(e, Tree) and e.data == "functioncall":
            return self.eval_call(e, sc)
        if isinstance(e, Token) and e.type == "VARARG":
            s = sc.lookup("...")
            return list(s.vars["..."]) if s else []
        return [self.eval(e, sc)]

    def eval_call(self, node, sc):
        kids = node.children
        if len(kids) == 2:
            f = self.eval(kids[0], sc)
            args = self.eval_args(kids[1], sc)
        else:
            obj = self.eval(kids[0], sc)
            f = self.index(obj, str(kids[1]))
            if f is None:
                raise LuaError("attempt to call method '%s' (a nil value)" % kids[1])
            args = [obj] + self.eval_args(kids[2], sc)
        return self.call(f, args)

    def eval_args(self, args, sc):
        if not args.children:
            return []
        c = args.children[0]
        if isinstance(c, Tree) and c.data == "explist":
            return self.eval_list(c, sc)
        return [self.eval(c, sc)]

    def eval(self, e, sc):
        if isinstance(e, Token):
            t = e.type
            if t == "NUMBER":
                s = str(e)
                if s.lower().startswith("0x"):
                    return int(s, 16)
                f = float(s)
                return int(f) if f.is_integer() and "." not in s and "e" not in s.lower() else f
            if t == "STRING":
                return bytes(str(e)[1:-1], "utf-8").decode("unicode_escape")
            if t == "LONGSTRING":
                s = str(e)
                lvl = s.index("[", 1) + 1
                body = s[lvl:len(s) - lvl]
                return body[1:] if body.startswith("\n") else body
            if t == "NIL":
                return None
            if t == "TRUE":
                return True
            if t == "FALSE":
                return False
            if t == "VARARG":
                v = self.eval_multi(e, sc)
                return v[0] if v else None
            if t == "NAME":
                return self.lookup_name(str(e), sc)
            raise LuaError("token %s" % t)
        d = e.data
        if d == "var":
            kids = e.children
            if len(kids) == 1:
                return self.lookup_name(str(kids[0]), sc)
            obj = self.eval(kids[0], sc)
            key = str(kids[1]) if isinstance(kids[1], Token) and kids[1].type == "NAME" else self.eval(kids[1], sc)
            return self.index(obj, key)
        if d == "functioncall":
            r = self.eval_call(e, sc)
            return r[0] if r else None
        if d == "paren_exp":                          # (f()) keeps the first value only
            return self.eval(e.children[0], sc)
        if d == "function":
            return LuaFunction(self, e.children[0], sc, False)
        if d == "tableconstructor":
            t = LuaTable()
            n = 0
            fields = e.children[0].children if e.children else []
            for i, f in enumerate(fields):
                k = f.children
                if len(k) == 2 and isinstance(k[0], Token) and k[0].type == "NAME":
                    self.setindex(t, str(k[0]), self.eval(k[1], sc))
                elif len(k) == 2:
                    self.setindex(t, self.eval(k[0], sc), self.eval(k[1], sc))
                else:
                    vals = self.eval_multi(k[0], sc) if i == len(fields) - 1 else [self.eval(k[0], sc)]
                    for v in vals:
                        n += 1
                        if v is not None:
                            t.hash[n] = v
            return t
        if d == "or_exp":
            v = None
            for c in e.children:
                v = self.eval(c, sc)
                if self.truthy(v):
                    return v
            return v
        if d == "and_exp":
            v = None
            for c in e.children:
                v = self.eval(c, sc)
                if not self.truthy(v):
                    return v
            return v
        if d == "cmp_exp":
            kids = e.children
            a = self.eval(kids[0], sc)
            i = 1
            while i < len(kids):
                op, b = str(kids[i]), self.eval(kids[i + 1], sc)
                a = {"==": lambda x, y: x is y if isinstance(x, (LuaTable, LuaFunction)) or isinstance(y, (LuaTable, LuaFunction)) else x == y,
                     "~=": lambda x, y: not (x is y if isinstance(x, (LuaTable, LuaFunction)) or isinstance(y, (LuaTable, LuaFunction)) else x == y),
                     "<": lambda x, y: x < y, "<=": lambda x, y: x <= y, ">": lambda x, y: x > y, ">=": lambda x, y: x >= y}[op](a, b)
                i += 2
            return a
        if d == "cat_exp":
            return self.tostring(self.eval(e.children[0], sc)) + self.tostring(self.eval(e.children[1], sc))
        if d in ("add_exp", "mul_exp"):
            kids = e.children
            a = self.eval(kids[0], sc)
            i = 1
            while i < len(kids):
                op, b = str(kids[i]), self.eval(kids[i + 1], sc)
                if a is None or b is None:
                    raise LuaError("attempt to perform arithmetic on a nil value")
                if isinstance(a, LuaTable) or isinstance(b, LuaTable):          # arithmetic metamethods
                    event = {"+": "__add", "-": "__sub", "*": "__mul", "/": "__div", "%": "__mod"}[op]
                    h = None
                    for operand in (a, b):
                        if isinstance(operand, LuaTable) and operand.meta is not None and operand.meta.hash.get(event) is not None:
                            h = operand.meta.hash[event]
                            break
                    if h is None:
                        raise LuaError("attempt to perform arithmetic on a table value")
                    a = (self.call(h, [a, b]) or [None])[0]
                    i += 2
                    continue
                a = {"+": lambda x, y: x + y, "-": lambda x, y: x - y, "*": lambda x, y: x * y, "/": lambda x, y: x / y,
                     "%": lambda x, y: x - (x // y) * y}[op](a, b)
                i += 2
            return a
        if d == "unary_exp":
            op, v = str(e.children[0]), self.eval(e.children[1], sc)
            if op == "not":
                return not self.truthy(v)
            if op == "#":
                return v.length() if isinstance(v, LuaTable) else len(v)
            if isinstance(v, LuaTable):
                h = v.meta.hash.get("__unm") if v.meta is not None else None
                if h is None:
                    raise LuaError("attempt to perform arithmetic on a table value")
                return (self.call(h, [v, v]) or [None])[0]
            return -v
        if d == "pow_exp":
            return self.eval(e.children[0], sc) ** self.eval(e.children[1], sc)
        raise LuaError("expression not supported by the test interpreter: %s" % d)

    def lookup_name(self, name, sc):
        s = sc.lookup(name)
        return s.vars[name] if s else None


def to_lua(v):
    """Python list / dict -> LuaTable (recursively); other values unchanged."""
    if isinstance(v, list):
        return LuaTable({i + 1: to_lua(x) for i, x in enumerate(v)})
    if isinstance(v, dict):
        return LuaTable({k: to_lua(x) for k, x in v.items()})
    return v
