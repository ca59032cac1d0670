"""The REAL reference core (radio/core/class.lua, util.lua, block.lua, composite.lua) and the REAL block files of the hot
path, loaded from /root/reference into the test interpreter (tests/lua_interp.py), with mocks only for what needs LuaJIT's
FFI or the OS: `ffi`, `radio.core.pipe`, `radio.core.debug`, `radio.core.platform`, `radio.types`, `radio.core.vector`.
Used by tests/test_lua_reference.py to run the glue in lua/radio_b200/ against the reference's own class system and
CompositeBlock:_prepare_to_run().  Nothing here is available on a machine without the reference tree (the tests skip)."""
import os

from tests.lua_interp import Interp, LuaTable

REF = os.environ.get("LUARADIO_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LUA = os.path.join(ROOT, "lua", "radio_b200")

# reference class name -> module path (real files)
REAL_BLOCKS = {
    "FIRFilterBlock": "blocks/signal/firfilter", "LowpassFilterBlock": "blocks/signal/lowpassfilter",
    "HighpassFilterBlock": "blocks/signal/highpassfilter", "BandpassFilterBlock": "blocks/signal/bandpassfilter",
    "ComplexBandpassFilterBlock": "blocks/signal/complexbandpassfilter",
    "FrequencyTranslatorBlock": "blocks/signal/frequencytranslator", "FrequencyDiscriminatorBlock": "blocks/signal/frequencydiscriminator",
    "DownsamplerBlock": "blocks/signal/downsampler", "UpsamplerBlock": "blocks/signal/upsampler",
    "IIRFilterBlock": "blocks/signal/iirfilter", "SinglepoleLowpassFilterBlock": "blocks/signal/singlepolelowpassfilter",
    "SinglepoleHighpassFilterBlock": "blocks/signal/singlepolehighpassfilter", "FMDeemphasisFilterBlock": "blocks/signal/fmdeemphasisfilter",
    "HilbertTransformBlock": "blocks/signal/hilberttransform", "ComplexMagnitudeBlock": "blocks/signal/complexmagnitude",
    "ComplexToRealBlock": "blocks/signal/complextoreal", "MultiplyConstantBlock": "blocks/signal/multiplyconstant",
    "MultiplyBlock": "blocks/signal/multiply", "MultiplyConjugateBlock": "blocks/signal/multiplyconjugate", "AddBlock": "blocks/signal/add",
    "SubtractBlock": "blocks/signal/subtract", "DelayBlock": "blocks/signal/delay", "PLLBlock": "blocks/signal/pll",
    "IQFileSource": "blocks/sources/iqfile", "RealFileSource": "blocks/sources/realfile", "IQFileSink": "blocks/sinks/iqfile",
    "RealFileSink": "blocks/sinks/realfile",
}
REAL_COMPOSITES = {
    "TunerBlock": "composites/tuner", "DecimatorBlock": "composites/decimator", "WBFMMonoDemodulator": "composites/wbfmmonodemodulator",
    "WBFMStereoDemodulator": "composites/wbfmstereodemodulator", "AMSynchronousDemodulator": "composites/amsynchronousdemodulator",
    "NBFMDemodulator": "composites/nbfmdemodulator", "AMEnvelopeDemodulator": "composites/amenvelopedemodulator",
    "SSBDemodulator": "composites/ssbdemodulator", "InterpolatorBlock": "composites/interpolator",
    "RationalResamplerBlock": "composites/rationalresampler",
}
REAL_UTILS = ["utilities/filter_utils", "utilities/window_utils", "utilities/format_utils", "utilities/spectrum_utils"]


def available():
    return os.path.isfile(os.path.join(REF, "radio", "core", "composite.lua"))


def _read(rel):
    with open(os.path.join(REF, "radio", rel + ".lua")) as f:
        return f.read()


def _glue(name):
    with open(os.path.join(LUA, name + ".lua")) as f:
        return f.read()


def make_env(lib, cuda=True):
    """Interpreter with the reference core + hot-path blocks + the glue registered as modules; `lib` is the mock library
    platform.load() hands out.  Returns (interp, types)."""
    vector_class = LuaTable({"name": "Vector"})

    def make_type(name):
        t = LuaTable({"type_name": name})

        def vector(n=0):
            v = LuaTable({"data": LuaTable(), "length": n or 0, "data_type": t, "_types": LuaTable({vector_class: True})})
            for i in range(int(n or 0)):
                v.hash["data"].hash[i] = LuaTable({"value": 0.0, "real": 0.0, "imag": 0.0})

            def resize(self, m):
                self.hash["length"] = m
                return [self]
            v.hash["resize"] = resize
            return [v]

        def vector_from_array(arr):
            v = vector(arr.length())[0]
            for i, x in enumerate(arr.array()):
                v.hash["data"].hash[i] = x if isinstance(x, LuaTable) else LuaTable({"value": x})
            return [v]
        t.hash["vector"] = vector
        t.hash["vector_from_array"] = vector_from_array
        return t

    types = LuaTable({k: make_type(k) for k in ("ComplexFloat32", "Float32", "Bit", "Byte")})
    # ComplexFloat32(re, im) is called by a few blocks at module level / initialize
    types.hash["ComplexFloat32"].meta = LuaTable({"__call": lambda self, re=0.0, im=0.0: [LuaTable({"real": re, "imag": im})]})
    types.hash["Float32"].meta = LuaTable({"__call": lambda self, v=0.0: [LuaTable({"value": v})]})
    ffi = LuaTable({
        "cdef": lambda text: [], "typeof": lambda *a: [LuaTable()], "new": lambda ct, *a: [LuaTable({0: 0, "ctype": ct})],
        "metatype": lambda ct, mt: [mt], "C": LuaTable({k: (lambda *a: [0]) for k in ("sigemptyset", "sigaddset", "sigprocmask", "signal", "sigpending",
                                                                                  "sigismember", "getpid", "kill", "waitpid")}), "string": lambda s, *a: [s], "gc": lambda o, f: [o], "sizeof": lambda *a: [8],
        "cast": lambda t, v: [v], "istype": lambda ct, v: [isinstance(v, LuaTable) and v.hash.get("ctype") is ct], "errno": lambda: [0],
        "copy": lambda *a: [], "fill": lambda *a: [], "abi": lambda what: [what == "le"],
    })

    def new_pipe(cls, output, input_):
        p = LuaTable({"output": output, "input": input_, "initialized": False})
        p.hash["get_rate"] = lambda self: it.call(it.index(output.hash["owner"], "get_rate"), [output.hash["owner"]])

        def initialize(self, *a):
            self.hash["initialized"] = True
            return []
        p.hash["initialize"] = initialize
        p.hash["write"] = lambda self, vec: self.hash.setdefault("queue", []).append(vec) or []
        return [p]

    def new_socket(cls):
        return [LuaTable({"initialize": lambda self: []})]

    # PipeMux for the single-process run loop (block.lua:493-556): one vector per input per run_once(), EOF when an input
    # queue is empty; write() appends to every downstream pipe's queue
    def new_mux(cls, input_pipes, output_pipes, control_socket=None):
        def read(self):
            data = LuaTable()
            for i, p in enumerate(input_pipes.array(), 1):
                q = p.hash.setdefault("queue", [])
                if not q:
                    return [data, True, False]
                data.hash[i] = q.pop(0)
            return [data, False, False]

        def write(self, data_out):
            for i, pipes in enumerate(output_pipes.array(), 1):
                for p in pipes.array():
                    p.hash.setdefault("queue", []).append(data_out.hash[i])
            return [False, None, False]
        return [LuaTable({"read": read, "write": write})]
    pipe = LuaTable({"Pipe": LuaTable(), "ControlSocket": LuaTable(), "PipeMux": LuaTable()})
    pipe.hash["Pipe"].meta = LuaTable({"__call": new_pipe})
    pipe.hash["ControlSocket"].meta = LuaTable({"__call": new_socket})
    pipe.hash["PipeMux"].meta = LuaTable({"__call": new_mux})
    debug = LuaTable({"print": lambda *a: [], "printf": lambda *a: [], "enabled": False})
    platform = LuaTable({"features": LuaTable(), "libs": LuaTable(), "os": "Linux", "arch": "x64", "page_size": 4096,
                         "load": lambda names: [cuda, lib if cuda else None], "alloc": lambda n: [LuaTable()]})
    modules = {"ffi": ffi, "radio.core.pipe": pipe, "radio.core.debug": debug, "radio.core.platform": platform, "radio.types": types,
               "radio.core.vector": LuaTable({"Vector": vector_class, "ObjectVector": LuaTable()}),
               "radio.core.async": LuaTable({"callback": lambda *a: [None]})}
    for rel in ("core/class", "core/util", "core/block", "core/composite"):
        modules["radio." + rel.replace("/", ".")] = _read(rel)
    for rel in list(REAL_BLOCKS.values()) + list(REAL_COMPOSITES.values()) + REAL_UTILS:
        modules["radio." + rel.replace("/", ".")] = _read(rel)
    for f in os.listdir(LUA):
        if f.endswith(".lua"):
            modules["radio_b200." + f[:-4]] = _glue(f[:-4])
            if f == "init.lua":
                modules["radio_b200"] = modules["radio_b200.init"]
    # `radio` / `radio.blocks`: the registry the composites and user scripts see -- only the hot-path classes
    lines = ["local radio = {}", "package_loaded_radio = radio",
             "radio.CompositeBlock = require('radio.core.composite').CompositeBlock",
             "radio.block = require('radio.core.block')", "radio.types = require('radio.types')"]
    for name, rel in REAL_BLOCKS.items():
        lines.append("radio.%s = require('radio.%s')" % (name, rel.replace("/", ".")))
    lines.append("return radio")
    modules["radio.blocks"] = "\n".join(lines)
    comp = ["local radio = require('radio.blocks')"]
    for name, rel in REAL_COMPOSITES.items():
        comp.append("radio.%s = require('radio.%s')" % (name, rel.replace("/", ".")))
    comp.append("return radio")
    modules["radio"] = "\n".join(comp)
    it = Interp(modules)
    it.f32 = Float32Types(it, types, vector_class)
    ffi.hash["sizeof"] = it.f32.sizeof
    ffi.hash["copy"] = lambda dst, src, n=None: it.f32.memmove(dst, src, n) and []
    ffi.hash["C"].hash["memmove"] = it.f32.memmove
    it.modules["string"] = it.G.vars["string"]
    it.modules["math"] = it.G.vars["math"]
    it.modules["io"] = it.G.vars["io"]
    it.modules["os"] = it.G.vars["os"]
    return it, types


class Float32Types:
    """Float32-faithful stand-ins for the reference's cdata sample types, for RUNNING the reference's pure-Lua process() loops
    in the interpreter: complex_float32_t {float real, imag} and float32_t {float value} are cells that round to float32 on
    every store (radio/types/complexfloat32.lua:19-24, float32.lua:17-21); ComplexFloat32's operators and methods follow
    complexfloat32.lua (products and sums are computed on Lua numbers = doubles and rounded when the result is constructed;
    arg() is atan2f, abs() sqrtf).  Vectors made by make_vector() create their cells on first access."""

    def __init__(self, it, types, vector_class):
        import numpy as np
        self.np, self.it, self.types, self.vector_class = np, it, types, vector_class
        f32 = lambda v: float(np.float32(v))
        self.f32 = f32

        def cell_meta(fields):
            store = {}

            def index(t, k):
                return [store.get((id(t), k), 0.0 if k in fields else None)]

            def newindex(t, k, v):
                store[(id(t), k)] = f32(v) if k in fields else v
                return []
            return index, newindex, store
        # ---- ComplexFloat32
        cidx, cnew, cstore = cell_meta(("real", "imag"))
        self._cstore = cstore
        methods = {}

        def cindex(t, k):
            if k in methods:
                return [methods[k]]
            return cidx(t, k)
        self.cmeta = LuaTable({"__index": cindex, "__newindex": cnew})

        def cnew_value(re=0.0, im=0.0):
            t = LuaTable()
            t.meta = self.cmeta
            cstore[(id(t), "real")] = f32(re)
            cstore[(id(t), "imag")] = f32(im)
            self._keep.append(t)
            return t
        self._keep = []
        self._where, self._elem_size, self._is_complex, self._cells = {}, {}, {}, {}     # cell -> (data table, index); data table -> element size / kind; (data table, index) -> cell
        self.complex = cnew_value
        re = lambda t: cstore[(id(t), "real")]
        im = lambda t: cstore[(id(t), "imag")]
        self.re, self.im = re, im
        self.cmeta.hash["__mul"] = lambda a, b: [cnew_value(re(a) * re(b) - im(a) * im(b), re(a) * im(b) + im(a) * re(b))]
        self.cmeta.hash["__add"] = lambda a, b: [cnew_value(re(a) + re(b), im(a) + im(b))]
        self.cmeta.hash["__sub"] = lambda a, b: [cnew_value(re(a) - re(b), im(a) - im(b))]
        methods["conj"] = lambda a: [cnew_value(re(a), -im(a))]
        methods["arg"] = lambda a: [float(np.arctan2(np.float32(im(a)), np.float32(re(a)), dtype=np.float32))]
        methods["abs"] = lambda a: [float(np.sqrt(np.float32(re(a) * re(a) + im(a) * im(a)), dtype=np.float32))]
        methods["abs_squared"] = lambda a: [re(a) * re(a) + im(a) * im(a)]
        methods["scalar_mul"] = lambda a, k: [cnew_value(re(a) * k, im(a) * k)]
        methods["scalar_div"] = lambda a, k: [cnew_value(re(a) / k, im(a) / k)]
        # ---- Float32
        fidx, fnew, fstore = cell_meta(("value",))
        self._fstore = fstore
        self.fmeta = LuaTable({"__index": fidx, "__newindex": fnew})

        def fnew_value(v=0.0):
            t = LuaTable()
            t.meta = self.fmeta
            fstore[(id(t), "value")] = f32(v)
            self._keep.append(t)
            return t
        self.real = fnew_value
        val = lambda t: fstore[(id(t), "value")]
        self.fmeta.hash["__add"] = lambda a, b: [fnew_value(val(a) + val(b))]            # radio/types/float32.lua:51-85
        self.fmeta.hash["__sub"] = lambda a, b: [fnew_value(val(a) - val(b))]
        self.fmeta.hash["__mul"] = lambda a, b: [fnew_value(val(a) * val(b))]
        self.fmeta.hash["__div"] = lambda a, b: [fnew_value(val(a) / val(b))]
        C, F = types.hash["ComplexFloat32"], types.hash["Float32"]
        C.meta = LuaTable({"__call": lambda self_, r=0.0, i=0.0: [cnew_value(r, i)]})
        F.meta = LuaTable({"__call": lambda self_, v=0.0: [fnew_value(v)]})
        C.hash["vector"] = lambda n=0: [self.make_vector(C, n)]
        F.hash["vector"] = lambda n=0: [self.make_vector(F, n)]

        def from_array(data_type):
            def f(arr):
                v = self.make_vector(data_type, arr.length())
                for i, x in enumerate(arr.array()):
                    self.set_cell(v.hash["data"], i, x if isinstance(x, LuaTable) else (cnew_value(x, 0.0) if data_type is C else fnew_value(x)))
                return [v]
            return f
        C.hash["vector_from_array"] = from_array(C)
        F.hash["vector_from_array"] = from_array(F)

    def make_vector(self, data_type, n=0):
        is_c = data_type is self.types.hash["ComplexFloat32"]
        data = LuaTable()                           # a PROXY: its own hash stays empty, so __index / __newindex always fire
        cells = self._cells

        def create(t, k):
            cell = cells.get((id(t), k))
            if cell is None:
                cell = self.complex() if is_c else self.real()
                cells[(id(t), k)] = cell
                self._where[id(cell)] = (t, k)
            return [cell]

        def assign(t, k, v):                       # out.data[i] = <value>: a struct copy in LuaJIT, every time
            if is_c:
                if v.meta is self.cmeta:
                    cell = self.complex(self.re(v), self.im(v))
                else:                              # a {re, im} pair (vector_from_array of filter_utils' complex taps)
                    cell = self.complex(v.hash.get(1, v.hash.get("real", 0.0)), v.hash.get(2, v.hash.get("imag", 0.0)))
            else:
                cell = self.real(v if not isinstance(v, LuaTable) else self._fstore[(id(v), "value")])
            cells[(id(t), k)] = cell
            self._where[id(cell)] = (t, k)
            return []
        self._elem_size[id(data)] = 8 if is_c else 4
        self._is_complex[id(data)] = is_c
        self._keep.append(data)
        data.meta = LuaTable({"__index": create, "__newindex": assign})
        v = LuaTable({"data": data, "length": int(n or 0), "data_type": data_type, "_types": LuaTable({self.vector_class: True})})

        def resize(self_, m):
            self_.hash["length"] = int(m)
            return [self_]
        v.hash["resize"] = resize
        return v

    def set_cell(self, data, k, value):
        self.it.call(data.meta.hash["__newindex"], [data, k, value])

    def get_cell(self, data, k):
        return self.it.call(data.meta.hash["__index"], [data, k])[0]

    def vector_from_numpy(self, x):
        np = self.np
        is_c = np.iscomplexobj(x)
        v = self.make_vector(self.types.hash["ComplexFloat32" if is_c else "Float32"], len(x))
        for i, s in enumerate(x):
            self.set_cell(v.hash["data"], i, self.complex(s.real, s.imag) if is_c else self.real(float(s)))
        return v

    def to_numpy(self, v):
        np = self.np
        n = v.hash["length"]
        d = v.hash["data"]
        if v.hash["data_type"] is self.types.hash["ComplexFloat32"]:
            return np.array([complex(self.re(self.get_cell(d, i)), self.im(self.get_cell(d, i))) for i in range(n)], np.complex64)
        return np.array([self._fstore[(id(self.get_cell(d, i)), "value")] for i in range(n)], np.float32)

    # ---- pointer emulation for the reference's pure-Lua branches: `v.data` is element 0, `v.data[k]` passed to memmove /
    # ffi.copy is the address of element k (LuaJIT passes a struct reference where a pointer is expected)
    def _locate(self, p):
        if id(p) in self._elem_size:
            return p, 0
        if id(p) in self._where:
            return self._where[id(p)]
        from tests.lua_interp import LuaError
        raise LuaError("memmove / ffi.copy: not a vector position")

    def sizeof(self, p, *a):
        if isinstance(p, LuaTable):
            if p.meta is self.cmeta:
                return [8]
            if p.meta is self.fmeta:
                return [4]
        return [8]

    def memmove(self, dst, src, nbytes):
        (dt, di), (st, si) = self._locate(dst), self._locate(src)
        size = self._elem_size[id(st)]
        assert self._elem_size[id(dt)] == size and nbytes % size == 0
        n = int(nbytes // size)
        is_c = self._is_complex[id(st)]
        vals = []
        for k in range(n):
            cell = self._cells.get((id(st), si + k))
            if cell is None:
                vals.append((0.0, 0.0) if is_c else 0.0)
            else:
                vals.append((self.re(cell), self.im(cell)) if is_c else self._fstore[(id(cell), "value")])
        for k, v in enumerate(vals):
            cell = self.complex(*v) if is_c else self.real(v)
            self._cells[(id(dt), di + k)] = cell
            self._where[id(cell)] = (dt, di + k)
        return [dst]
