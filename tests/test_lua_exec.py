"""The LuaJIT glue (lua/radio_b200/*.lua) EXECUTED under the test interpreter of tests/lua_interp.py, against a mock `radio`
class system (tests/lua_mock/), a mock `ffi` and a mock libluaradio_b200 that records every call and only answers for
symbols declared in the generated cdef.  LuaJIT itself is not installed in the build image; this is how the logic of the
glue -- the scheduler's collapse_gpu_runs, the class patching, the per-block create calls and their arguments, the
process()/cleanup() call sequences -- gets run in the CPU suite.  The scheduler's result is compared with the Python planner
(CompositeBlock._plan_gpu_runs) on the same flow graphs."""
import os
import re

import numpy as np
import pytest

from tests.lua_interp import Interp, LuaError, LuaTable, to_lua

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LUA = os.path.join(ROOT, "lua", "radio_b200")
MOCK = os.path.join(ROOT, "tests", "lua_mock")


def read(path):
    with open(path) as f:
        return f.read()


class Handle:
    def __init__(self, what, args):
        self.what, self.args = what, args

    def __repr__(self):
        return "<%s>" % self.what


class MockLib:
    """libluaradio_b200 as the glue sees it through ffi.load: only declared symbols resolve (LuaJIT raises "missing
    declaration" otherwise); every call is logged as (name, args)."""

    def __init__(self, declared):
        self._declared, self.calls, self.fail_next_create, self.dag_nodes = declared, [], False, 0

    def __getattr__(self, name):
        if name.startswith("_") or name in ("calls", "fail_next_create", "dag_nodes"):
            raise AttributeError(name)
        if name not in self._declared:
            raise LuaError("missing declaration for symbol '%s'" % name)

        def fn(*args):
            self.calls.append((name, args))
            if name.endswith("_create") or "_create_" in name:
                if self.fail_next_create:
                    self.fail_next_create = False
                    return [None]
                return [Handle(name, args)]
            if name == "lrb200_device_count":
                return [1]
            if name == "lrb200_last_error":
                return ["mock error text"]
            if name.endswith("max_output"):
                return [args[-1]]
            if name in ("lrb200_dag_add_graph", "lrb200_dag_add_block"):
                self.dag_nodes += 1
                return [self.dag_nodes - 1]
            if name == "lrb200_dag_execute":
                for k in list(args[4].hash):
                    if isinstance(k, int):
                        args[4].hash[k] = args[2]
                args[4].hash[0] = args[2]
                args[4].hash[1] = args[2]
                return [0]
            if name == "lrb200_block_execute_multi":
                args[6].hash[0] = args[3]
                return [0]
            if name in ("lrb200_block_execute", "lrb200_graph_execute"):
                args[4].hash[0] = args[2]
                return [0]
            if name == "lrb200_graph_flush":
                args[2].hash[0] = 3
                return [0]
            return [0]
        return fn


def make_env(cuda=True):
    """An interpreter with the mock modules registered; returns (interp, lib, types)."""
    cdef_mod = Interp().run(read(os.path.join(LUA, "cdef.lua")))[0]
    declared = set(re.findall(r"\b(lrb200_\w+)\s*\(", cdef_mod.hash["cdef"]))
    lib = MockLib(declared)

    def make_type(name):
        t = LuaTable({"type_name": name})

        def vector(n=0):
            v = LuaTable({"data": "%s*" % name, "length": n, "data_type": t})

            def resize(self, m):
                self.hash["length"] = m
                return [self]
            v.hash["resize"] = resize
            return [v]
        t.hash["vector"] = vector
        return t

    types = LuaTable({"ComplexFloat32": make_type("ComplexFloat32"), "Float32": make_type("Float32"), "Bit": make_type("Bit")})
    cdefs = []
    ffi_c = LuaTable({"fwrite": lambda data, size, n, f: [n], "fread": lambda data, size, n, f: [n], "feof": lambda f: [0]})
    ffi = LuaTable({
        "cdef": lambda text: cdefs.append(text) or [],
        "gc": lambda obj, fin: [obj],
        "new": lambda ct, *a: [LuaTable({0: 0, "ctype": ct})],
        "string": lambda s: [s],
        "istype": lambda ct, v: [isinstance(v, LuaTable) and v.hash.get("ctype") is ct],
        "sizeof": lambda ct: [8],
        "C": ffi_c,
    })
    platform = LuaTable({"features": LuaTable(), "libs": LuaTable(), "load": lambda names: [cuda, lib if cuda else None]})
    modules = {
        "ffi": ffi, "radio.core.platform": platform, "radio.types": types,
        "radio.core.block": read(os.path.join(MOCK, "block.lua")), "radio": read(os.path.join(MOCK, "radio.lua")),
    }
    for f in os.listdir(LUA):
        if f.endswith(".lua"):
            modules["radio_b200." + f[:-4]] = read(os.path.join(LUA, f))
            if f == "init.lua":
                modules["radio_b200"] = modules["radio_b200.init"]
    it = Interp(modules)
    it.cdefs = cdefs
    return it, lib, types


def test_interpreter_basics():
    """The interpreter itself: closures, multiple returns, metatables, varargs, numeric/generic for, and the pairs() rule."""
    it = Interp()
    r = it.run("""
        local function f(a, ...) local t = {...} return a + #t, select('#', ...) end
        local C = {} C.__index = C
        function C:get() return self.v end
        local o = setmetatable({v = 7}, C)
        local acc = 0
        for i = 1, 10, 3 do acc = acc + i end            -- 1 + 4 + 7 + 10
        local keys = {}
        for k, v in pairs({a = 1, b = 2}) do keys[#keys + 1] = k .. v end
        table.insert(keys, 1, "first")
        local n = 0
        repeat local done = n >= 2; n = n + 1 until done
        local x, y = f(1, 'p', 'q')
        return x, y, o:get(), acc, #keys, keys[1], n, (nil or false) == false, 7 % 3, 2 ^ 10, "a" .. 1
    """)
    assert r == [3, 2, 7, 22, 3, "first", 3, True, 1, 1024, "a1"]
    with pytest.raises(LuaError, match="pairs"):
        it.run("local t = {a = 1} for k in pairs(t) do t[k .. 'x'] = 1 end")
    it.run("local t = {a = 1, b = 2} for k in pairs(t) do t[k] = nil end")       # clearing is allowed
    with pytest.raises(LuaError, match="index a nil"):
        it.run("local t = {} return t.a.b")


def test_platform_probe_sets_the_feature_and_respects_the_kill_switch(monkeypatch):
    monkeypatch.delenv("LUARADIO_DISABLE_CUDA", raising=False)
    monkeypatch.setenv("LUARADIO_CUDA_DEVICE", "3")
    it, lib, _ = make_env()
    b200 = it.require("radio_b200.platform")
    plat = it.require("radio.core.platform")
    assert plat.hash["features"].hash["cuda"] is True and plat.hash["libs"].hash["cuda"] is lib
    assert [c for c in lib.calls] == [("lrb200_device_count", ()), ("lrb200_init", (3,))]
    assert len(it.cdefs) == 1 and "lrb200_graph_execute" in it.cdefs[0]
    assert (b200.hash["HOST"], b200.hash["DEVICE"], b200.hash["FIR_DIRECT"], b200.hash["FIR_FFT"]) == (0, 1, 1, 2)
    with pytest.raises(LuaError, match="graph_commit: mock error text"):
        it.call(b200.hash["fail"], ["graph_commit"])
    # kill switch, and a box where the library does not load
    monkeypatch.setenv("LUARADIO_DISABLE_CUDA", "1")
    it, lib, _ = make_env()
    it.require("radio_b200.platform")
    assert it.require("radio.core.platform").hash["features"].hash["cuda"] is False and lib.calls == []
    monkeypatch.delenv("LUARADIO_DISABLE_CUDA")
    it, lib, _ = make_env(cuda=False)
    it.require("radio_b200.platform")
    assert it.require("radio.core.platform").hash["features"].hash["cuda"] is False


def patched_radio(monkeypatch):
    monkeypatch.delenv("LUARADIO_DISABLE_CUDA", raising=False)
    monkeypatch.delenv("LUARADIO_CUDA_DEVICE", raising=False)
    monkeypatch.delenv("LUARADIO_B200_SUPERCHUNK", raising=False)
    it, lib, types = make_env()
    radio = it.require("radio")
    it.call(it.require("radio_b200.blocks_patch"), [radio])
    lib.calls.clear()
    return it, lib, types, radio


def vec(types, name, n):
    return types.hash[name].hash["vector"](n)[0]


def test_block_patches_create_the_right_handles(monkeypatch):
    """Every patched class: initialize() makes the HOST-pointer handle with the block's parameters, make_device_handle() the
    DEVICE one; process() is max_output -> execute -> resize; a NULL handle raises with the library's message."""
    it, lib, types, radio = patched_radio(monkeypatch)
    C, F = types.hash["ComplexFloat32"], types.hash["Float32"]
    new = lambda cls, *a: it.call(radio.hash[cls], list(a))[0]
    meth = lambda obj, name, *a: it.call(it.index(obj, name), [obj] + list(a))
    taps = vec(types, "Float32", 128)

    fir = new("FIRFilterBlock", taps, None, C)
    meth(fir, "initialize")
    assert lib.calls == [("lrb200_fir_create_crcf", (taps.hash["data"], 128, 1, 0))]        # use_fft nil: automatic, no set_algorithm
    assert fir.hash["out"].hash["data_type"] is C
    lib.calls.clear()
    fir = new("FIRFilterBlock", taps, True, F)
    h = meth(fir, "make_device_handle")[0]
    assert [c[0] for c in lib.calls] == ["lrb200_fir_create_rrrf", "lrb200_fir_set_algorithm"]
    assert lib.calls[0][1][3] == 1 and lib.calls[1][1] == (h, 2)                             # DEVICE pointers; FIR_FFT
    lib.calls.clear()
    ctaps = vec(types, "ComplexFloat32", 33)
    fir = new("FIRFilterBlock", ctaps, False, C)
    meth(fir, "initialize")
    assert [c[0] for c in lib.calls] == ["lrb200_fir_create_cccf", "lrb200_fir_set_algorithm"] and lib.calls[1][1][1] == 1
    # process(): the shared body
    lib.calls.clear()
    x = vec(types, "ComplexFloat32", 4096)
    y = it.call(it.index(fir, "process_complex_input_complex_taps"), [fir, x])[0]
    assert [c[0] for c in lib.calls] == ["lrb200_block_max_output", "lrb200_block_execute"]
    assert lib.calls[1][1][:4] == (fir.hash["handle"], x.hash["data"], 4096, y.hash["data"]) and y.hash["length"] == 4096
    for name in ("process_complex_input_real_taps", "process_real_input_real_taps", "process_fft_complex_input_complex_taps",
                 "process_fft_complex_input_real_taps", "process_fft_real_input_real_taps"):
        assert it.index(fir, name) is it.index(fir, "process_complex_input_complex_taps")

    # a subclass made AFTER the patch (Lowpass & co. only design taps) inherits the GPU form
    block = it.require("radio.core.block")
    lowpass = it.call(block.hash["factory"], ["LowpassFilterBlock", radio.hash["FIRFilterBlock"]])[0]
    assert it.index(lowpass, "make_device_handle") is it.index(radio.hash["FIRFilterBlock"], "make_device_handle")

    cases = [
        ("FrequencyTranslatorBlock", (-250e3,), {"rate": 1e6}, "lrb200_rotator_create", lambda a: a == (-0.25, 0)),
        ("FrequencyDiscriminatorBlock", (1.25,), {}, "lrb200_discrim_create", lambda a: a == (1.25, 0)),
        ("DownsamplerBlock", (5, C), {}, "lrb200_downsample_create", lambda a: a == (5, 8, 0)),
        ("DownsamplerBlock", (4, F), {}, "lrb200_downsample_create", lambda a: a == (4, 4, 0)),
        ("UpsamplerBlock", (3, F), {}, "lrb200_upsample_create", lambda a: a == (3, 4, 0)),
        ("IIRFilterBlock", (vec(types, "Float32", 2), vec(types, "Float32", 2), F), {}, "lrb200_iir_create_rrrf", lambda a: a[1] == 2 and a[3] == 2 and a[4] == 0),
        ("IIRFilterBlock", (vec(types, "Float32", 3), vec(types, "Float32", 3), C), {}, "lrb200_iir_create_crcf", lambda a: a[1] == 3 and a[4] == 0),
        ("HilbertTransformBlock", (vec(types, "Float32", 129),), {}, "lrb200_hilbert_create", lambda a: a[1:] == (129, 0)),
        ("ComplexMagnitudeBlock", (), {}, "lrb200_cmag_create", lambda a: a == (0,)),
        ("ComplexToRealBlock", (), {}, "lrb200_c2r_create", lambda a: a == (0,)),
        ("MultiplyConstantBlock", (LuaTable({"value": 2.5}), C), {}, "lrb200_mulconst_create", lambda a: a == (2.5, 0, 1, 0, 0)),
        ("MultiplyConstantBlock", (LuaTable({"real": 1.0, "imag": -1.0, "ctype": C}), C), {}, "lrb200_mulconst_create", lambda a: a == (1.0, -1.0, 1, 1, 0)),
    ]
    for cls, args, fields, symbol, check in cases:
        lib.calls.clear()
        b = new(cls, *args)
        for k, v in fields.items():
            b.hash[k] = v
        meth(b, "initialize")
        assert [c[0] for c in lib.calls] == [symbol], cls
        assert check(lib.calls[0][1]), (cls, lib.calls[0][1])
        assert isinstance(b.hash["handle"], Handle) and b.hash["out"] is not None
        lib.calls.clear()
        meth(b, "make_device_handle")
        assert lib.calls[0][0] == symbol and lib.calls[0][1][-1] == 1, cls
    # every class's process entry points are the shared body
    b200 = it.require("radio_b200.platform")
    for cls, names in (("FrequencyTranslatorBlock", ["process"]), ("IIRFilterBlock", ["process_complex", "process_real"]),
                       ("MultiplyConstantBlock", ["process", "process_complex_by_real"]), ("DownsamplerBlock", ["process"])):
        for n in names:
            assert radio.hash[cls].hash[n] is b200.hash["process"]
    # a NULL handle: error() with the library's text, like firfilter.lua's liquid branch
    lib.fail_next_create = True
    with pytest.raises(LuaError, match="Creating lrb200 cmag object: mock error text"):
        meth(new("ComplexMagnitudeBlock"), "initialize")
    lib.fail_next_create = True
    with pytest.raises(LuaError, match="Creating lrb200 cmag object"):
        meth(new("ComplexMagnitudeBlock"), "make_device_handle")


def test_file_blocks_remember_the_format_name_and_convert_through_the_library(monkeypatch):
    it, lib, types, radio = patched_radio(monkeypatch)
    meth = lambda obj, name, *a: it.call(it.index(obj, name), [obj] + list(a))
    sink = it.call(radio.hash["IQFileSink"], ["out.iq", "s16le", 1e6])[0]
    assert sink.hash["format_name"] == "s16le" and sink.hash["format"] == "s16le-table"       # the original instantiate() still ran
    sink.hash["raw_samples"] = vec(types, "Float32", 0)
    x = vec(types, "ComplexFloat32", 1000)
    meth(sink, "process", x)
    meth(sink, "process", x)
    assert [c[0] for c in lib.calls] == ["lrb200_iqsink_create", "lrb200_block_execute", "lrb200_block_execute"]   # one handle, reused
    assert lib.calls[0][1] == ("s16le", 0) and sink.hash["count"] == 2000 and sink.hash["raw_samples"].hash["length"] == 1000
    lib.calls.clear()
    src = it.call(radio.hash["RealFileSource"], ["in.f32", "f32le", 48e3])[0]
    src.hash["raw_samples"] = vec(types, "Float32", 8192)
    src.hash["out"] = vec(types, "Float32", 0)
    out = meth(src, "process")[0]
    assert lib.calls[0] == ("lrb200_realconv_create", ("f32le", 0)) and lib.calls[1][0] == "lrb200_block_execute"
    assert out.hash["length"] == 8192


def export_graph(it, radio, types, top, lua_gpu):
    """A prepared Python CompositeBlock's flat connection map, rebuilt from mock Lua blocks: same names, port counts,
    data types and edges.  Blocks whose class the Lua glue patches get that (patched) class, all others a host class."""
    from luaradio_b200.types import ComplexFloat32
    block = it.require("radio.core.block")
    lt = lambda t: types.hash["ComplexFloat32" if t is ComplexFloat32 else "Float32"]
    mk = it.run("""
        local block = require('radio.core.block')
        return function (class, name, ins, outs)
            local self = setmetatable({}, class)
            self.name = name
            local i, o = {}, {}
            for k, t in ipairs(ins) do i[k] = block.Input("in" .. k, t) end
            for k, t in ipairs(outs) do o[k] = block.Output("out" .. k, t) end
            self:add_type_signature(i, o)
            return self
        end
    """)[0]
    host = it.call(block.hash["factory"], ["HostOnly"])[0]
    lua_of, conns = {}, LuaTable()
    for b in top._concrete_order:
        cls = host
        if b in lua_gpu:
            base = radio.hash[lua_gpu[b]]
            cls = base if lua_gpu[b] == b.name else it.call(block.hash["factory"], [b.name, base])[0]
        lb = it.call(mk, [cls, b.name, to_lua([lt(p.data_type) for p in b.inputs]), to_lua([lt(p.data_type) for p in b.outputs])])[0]
        # the fields the create callbacks read
        for k, v in (("taps", vec(types, "Float32", 16)), ("offset", 1e5), ("rate", 1e6), ("gain", 1.0), ("factor", 2),
                     ("b_taps", vec(types, "Float32", 2)), ("a_taps", vec(types, "Float32", 2)), ("hilbert_taps", vec(types, "Float32", 9)),
                     ("constant", LuaTable({"value": 1.0})), ("num_samples", 129), ("loop_bw", 100.0), ("freq_min", 18950.0),
                     ("freq_max", 19050.0), ("multiplier", 2.0)):
            lb.hash[k] = v
        lua_of[b] = lb
    for inp, outp in top._all_connections.items():
        li = lua_of[inp.owner].hash["inputs"].hash[inp.owner.inputs.index(inp) + 1]
        lo = lua_of[outp.owner].hash["outputs"].hash[outp.owner.outputs.index(outp) + 1]
        conns.hash[li] = lo
    return lua_of, conns


# Python class name -> the Lua class that carries its GPU form (subclasses inherit through the metatable chain)
LUA_GPU_BASE = {
    "FIRFilterBlock": "FIRFilterBlock", "LowpassFilterBlock": "FIRFilterBlock", "HighpassFilterBlock": "FIRFilterBlock",
    "BandpassFilterBlock": "FIRFilterBlock", "BandstopFilterBlock": "FIRFilterBlock", "ComplexBandpassFilterBlock": "FIRFilterBlock",
    "ComplexBandstopFilterBlock": "FIRFilterBlock", "FrequencyTranslatorBlock": "FrequencyTranslatorBlock",
    "FrequencyDiscriminatorBlock": "FrequencyDiscriminatorBlock", "DownsamplerBlock": "DownsamplerBlock", "UpsamplerBlock": "UpsamplerBlock",
    "IIRFilterBlock": "IIRFilterBlock", "FMDeemphasisFilterBlock": "IIRFilterBlock", "SinglepoleLowpassFilterBlock": "IIRFilterBlock",
    "SinglepoleHighpassFilterBlock": "IIRFilterBlock", "HilbertTransformBlock": "HilbertTransformBlock",
    "ComplexMagnitudeBlock": "ComplexMagnitudeBlock", "ComplexToRealBlock": "ComplexToRealBlock", "MultiplyConstantBlock": "MultiplyConstantBlock",
    "MultiplyBlock": "MultiplyBlock", "MultiplyConjugateBlock": "MultiplyConjugateBlock", "AddBlock": "AddBlock", "SubtractBlock": "SubtractBlock",
    "DelayBlock": "DelayBlock", "PLLBlock": "PLLBlock",
}


def topologies():
    import luaradio_b200 as radio
    from luaradio_b200.block import Block, Input, Output
    from luaradio_b200.types import ComplexFloat32

    class Host(Block):
        name = "Host"

        def instantiate(self):
            self.add_type_signature([Input("in", ComplexFloat32)], [Output("out", ComplexFloat32)])

    x = np.zeros(16, np.complex64)
    tops = {}
    top = radio.CompositeBlock()
    top.connect(radio.ArraySource(x, 1102500.0), radio.TunerBlock(-250e3, 200e3, 5), radio.FrequencyDiscriminatorBlock(1.25),
                radio.LowpassFilterBlock(128, 15e3), radio.FMDeemphasisFilterBlock(75e-6), radio.DownsamplerBlock(5), radio.ArraySink())
    tops["mono"] = top
    top = radio.CompositeBlock()
    a1, a2, host, b1, b2, c1 = radio.FrequencyTranslatorBlock(1e5), radio.LowpassFilterBlock(64, 2e5), Host(), radio.DownsamplerBlock(2), \
        radio.ComplexMagnitudeBlock(), radio.ComplexToRealBlock()
    top.connect(radio.ArraySource(x, 1e6), a1, a2, host, b1)
    top.connect(b1, b2, radio.ArraySink())
    top.connect(b1, c1, radio.ArraySink())
    tops["host_and_fanout"] = top
    top = radio.CompositeBlock()
    mixer = radio.MultiplyConjugateBlock()
    top.connect(radio.ArraySource(x, 1e6), "out", mixer, "in1")
    top.connect(radio.ArraySource(x, 1e6), "out", mixer, "in2")
    top.connect(mixer, radio.LowpassFilterBlock(16, 100e3), radio.FrequencyDiscriminatorBlock(5), radio.DecimatorBlock(25, {"num_taps": 16}),
                radio.ArraySink())
    tops["two_sources"] = top
    top = radio.CompositeBlock()
    demod = radio.WBFMStereoDemodulator()
    top.connect(radio.ArraySource(x, 220500.0), demod)
    top.connect(demod, "left", radio.ArraySink(), "in")
    top.connect(demod, "right", radio.ArraySink(), "in")
    tops["stereo"] = top
    top = radio.CompositeBlock()
    top.connect(radio.ArraySource(x, 1e6), radio.AMEnvelopeDemodulator(), radio.ArraySink())
    tops["am_envelope"] = top
    for t in tops.values():
        t._prepare_to_run(initialize=False)
    return tops


@pytest.mark.parametrize("name", ["mono", "host_and_fanout", "two_sources", "stereo", "am_envelope"])
def test_lua_scheduler_matches_the_python_planner(monkeypatch, name):
    """collapse_gpu_runs (Lua, executed) == CompositeBlock._plan_gpu_runs (Python) on the same flattened graphs, and the
    rewritten connection map is consistent: members' edges gone, the chain wired to the run's old neighbours."""
    it, lib, types, radio = patched_radio(monkeypatch)
    top = topologies()[name]
    lua_gpu = {b: LUA_GPU_BASE[b.name] for b in top._concrete_order if b.name in LUA_GPU_BASE}
    host_side = {b for b in top._concrete_order if b not in lua_gpu}       # blocks without a Lua GPU form: host blocks for both planners
    expected = [[b.name for b in run] for run, _, _ in top._plan_gpu_runs(exclude=host_side) if len(run) >= 2]
    lua_of, conns = export_graph(it, radio, types, top, lua_gpu)
    before = dict(conns.hash)
    consumers_before = {}
    for i, o in before.items():
        consumers_before.setdefault(o, []).append(i)
    patch = it.require("radio_b200.composite_patch")
    out = it.call(patch.hash["collapse_gpu_runs"], [conns])[0]
    assert out is conns
    chains = {}
    for i, o in conns.hash.items():
        for port in (i, o):
            owner = port.hash["owner"]
            if "blocks" in owner.hash:
                chains[id(owner)] = owner
    got = [[b.hash["name"] for b in c.hash["blocks"].array()] for c in chains.values()]
    assert sorted(got) == sorted(expected), (got, expected)
    if name == "mono":
        assert got == [["FrequencyTranslatorBlock", "LowpassFilterBlock", "DownsamplerBlock", "FrequencyDiscriminatorBlock",
                        "LowpassFilterBlock", "FMDeemphasisFilterBlock", "DownsamplerBlock"]]
    if name == "stereo":
        assert ["FrequencyDiscriminatorBlock", "HilbertTransformBlock"] in got and got.count(["LowpassFilterBlock", "ComplexToRealBlock"]) == 2
    members = set()
    for c in chains.values():
        run = c.hash["blocks"].array()
        members.update(id(b) for b in run)
        first_in, last_out = run[0].hash["inputs"].hash[1], run[-1].hash["outputs"].hash[1]
        cin, cout = c.hash["inputs"].hash[1], c.hash["outputs"].hash[1]
        assert conns.hash[cin] is before[first_in]                                   # upstream edge re-targeted
        for consumer in consumers_before.get(last_out, []):
            assert conns.hash[consumer] is cout                                      # downstream consumers re-sourced
        assert cin.hash["data_type"] is first_in.hash["data_type"] and cout.hash["data_type"] is last_out.hash["data_type"]
        assert c.hash["differentiated"] is True
    for i, o in conns.hash.items():
        assert id(i.hash["owner"]) not in members and id(o.hash["owner"]) not in members
    # nothing else changed
    for i, o in before.items():
        if id(i.hash["owner"]) not in members and id(o.hash["owner"]) not in members:
            assert conns.hash[i] is o
    # every input port of every surviving block is still fed exactly once
    survivors = {id(p.hash["owner"]): p.hash["owner"] for pair in conns.hash.items() for p in pair}
    for b in survivors.values():
        for p in b.hash["inputs"].array():
            assert p in conns.hash


def test_chain_block_lifecycle_against_the_library(monkeypatch):
    """GPUChainBlock: initialize() = graph_create, one DEVICE-pointer handle per member appended in order, commit(fuse=1);
    process() = max_output -> execute; cleanup() = flush and a write to every downstream pipe; super-chunk from the env."""
    it, lib, types, radio = patched_radio(monkeypatch)
    top = topologies()["mono"]
    lua_gpu = {b: LUA_GPU_BASE[b.name] for b in top._concrete_order if b.name in LUA_GPU_BASE}
    lua_of, conns = export_graph(it, radio, types, top, lua_gpu)
    patch = it.require("radio_b200.composite_patch")
    it.call(patch.hash["collapse_gpu_runs"], [conns])
    chain = next(o.hash["owner"] for o in conns.hash.values() if "blocks" in o.hash["owner"].hash)
    meth = lambda obj, name, *a: it.call(it.index(obj, name), [obj] + list(a))
    assert meth(chain, "get_rate") == [1e6]                      # the last member's rate
    lib.calls.clear()
    meth(chain, "initialize")
    names = [c[0] for c in lib.calls]
    creates = ["lrb200_rotator_create", "lrb200_fir_create_crcf", "lrb200_downsample_create", "lrb200_discrim_create",
               "lrb200_fir_create_rrrf", "lrb200_iir_create_rrrf", "lrb200_downsample_create"]
    expect = ["lrb200_graph_create"]
    for c in creates:
        expect += [c, "lrb200_graph_append"]
    expect += ["lrb200_graph_commit"]
    assert names == expect
    graph = chain.hash["graph"]
    for k in range(len(creates)):
        create, append = lib.calls[1 + 2 * k], lib.calls[2 + 2 * k]
        assert create[1][-1] == 1                                # DEVICE pointers inside a flow graph
        assert append[1][0] is graph and isinstance(append[1][1], Handle) and append[1][1].what == creates[k]
    assert lib.calls[-1][1] == (graph, 1)
    assert lib.calls[5][1][:2] == (2, 8) and lib.calls[13][1][:2] == (2, 4)      # complex (8-byte) and real (4-byte) downsampler
    lib.calls.clear()
    x = vec(types, "ComplexFloat32", 8192)
    y = meth(chain, "process", x)[0]
    assert [c[0] for c in lib.calls] == ["lrb200_graph_max_output", "lrb200_graph_execute"]
    assert lib.calls[1][1][:4] == (graph, x.hash["data"], 8192, y.hash["data"]) and y is chain.hash["out"]
    # cleanup: the flushed samples go to every pipe of the output port
    written, turns = [], []
    consumer = LuaTable({"outputs": LuaTable(), "run_once": lambda self: turns.append(1) or [True]})
    pipe = LuaTable({"write": lambda self, v: written.append(v.hash["length"]) or [], "input": LuaTable({"owner": consumer})})
    chain.hash["outputs"].hash[1].hash["pipes"] = to_lua([pipe, pipe])
    chain.hash["evaluation_order"] = to_lua([chain, consumer])
    lib.calls.clear()
    meth(chain, "cleanup")
    assert [c[0] for c in lib.calls] == ["lrb200_graph_max_output", "lrb200_graph_flush"] and written == [3, 3]
    assert turns == [1]                                          # the block downstream takes one more turn on the flushed tail
    # the super-chunk switch
    monkeypatch.setenv("LUARADIO_B200_SUPERCHUNK", "1048576")
    lib.calls.clear()
    meth(chain, "initialize")
    assert lib.calls[-1] == ("lrb200_graph_set_superchunk", (chain.hash["graph"], 1048576))


def test_install_wraps_crawl_and_forces_one_process(monkeypatch):
    """install(radio): the TOP-LEVEL _crawl_connections call collapses the runs (recursive calls with an accumulator do not),
    and start() always reaches the original with multiprocess = false."""
    it, lib, types, radio = patched_radio(monkeypatch)
    top = topologies()["mono"]
    lua_gpu = {b: LUA_GPU_BASE[b.name] for b in top._concrete_order if b.name in LUA_GPU_BASE}
    lua_of, conns = export_graph(it, radio, types, top, lua_gpu)
    comp = it.call(radio.hash["CompositeBlock"], [])[0]
    comp.hash["flat"] = conns
    meth = lambda obj, name, *a: it.call(it.index(obj, name), [obj] + list(a))
    n_before = len(conns.hash)
    acc = LuaTable(dict(conns.hash))
    assert meth(comp, "_crawl_connections", acc, LuaTable())[0] is acc and len(acc.hash) == n_before      # nested call: untouched
    out = meth(comp, "_crawl_connections")[0]
    assert out is conns and len(conns.hash) == 2                 # source -> chain -> sink
    assert meth(comp, "start", True)[0] is comp and comp.hash["started_multiprocess"] is False
    assert meth(comp, "start")[0] is comp and comp.hash["started_multiprocess"] is False


def test_glue_is_inert_without_the_cuda_feature(monkeypatch):
    monkeypatch.setenv("LUARADIO_DISABLE_CUDA", "1")
    it, lib, types = make_env()
    radio = it.require("radio")
    before = dict(radio.hash["FIRFilterBlock"].hash)
    it.call(it.require("radio_b200.blocks_patch"), [radio])
    assert radio.hash["FIRFilterBlock"].hash == before and lib.calls == []
    assert "make_device_handle" not in radio.hash["DownsamplerBlock"].hash


def test_multi_port_block_patches(monkeypatch):
    """Multiply / MultiplyConjugate / Add / Subtract, DelayBlock and PLLBlock: create arguments, the execute_multi call of the
    shared multi-port process(), DelayBlock's host fallback for the Bit signature."""
    it, lib, types, radio = patched_radio(monkeypatch)
    C, F, B = types.hash["ComplexFloat32"], types.hash["Float32"], types.hash["Bit"]
    new = lambda cls, *a: it.call(radio.hash[cls], list(a))[0]
    meth = lambda obj, name, *a: it.call(it.index(obj, name), [obj] + list(a))
    b200 = it.require("radio_b200.platform")
    for cls, op in (("MultiplyBlock", "multiply"), ("MultiplyConjugateBlock", "multiplyconjugate"), ("AddBlock", "add"), ("SubtractBlock", "subtract")):
        for t, flag in ((C, 1), (F, 0)):
            lib.calls.clear()
            b = new(cls, t)
            meth(b, "initialize")
            assert lib.calls == [("lrb200_binary_create", (op, flag, 0))], cls
            assert b.hash["outs"].length() == 1 and b.hash["outs"].hash[1].hash["data_type"] is t
            lib.calls.clear()
            meth(b, "make_device_handle")
            assert lib.calls == [("lrb200_binary_create", (op, flag, 1))]
        assert radio.hash[cls].hash["process"] is b200.hash["process_multi"]
    assert radio.hash["MultiplyBlock"].hash["process_complex"] is b200.hash["process_multi"]
    assert radio.hash["MultiplyBlock"].hash["process_real"] is b200.hash["process_multi"]
    # process(x, y)
    b = new("MultiplyConjugateBlock", C)
    meth(b, "initialize")
    lib.calls.clear()
    x, y = vec(types, "ComplexFloat32", 512), vec(types, "ComplexFloat32", 512)
    y.hash["data"] = "second*"
    out = meth(b, "process", x, y)
    assert [c[0] for c in lib.calls] == ["lrb200_block_max_output", "lrb200_block_execute_multi"]
    h, ins, nin, n, outs, nout, _ = lib.calls[1][1]
    assert h is b.hash["handle"] and (nin, n, nout) == (2, 512, 1)
    assert (ins.hash[0], ins.hash[1]) == (x.hash["data"], "second*") and outs.hash[0] == out[0].hash["data"]
    assert len(out) == 1 and out[0].hash["length"] == 512
    # DelayBlock: float types on the GPU, Bit on the host
    lib.calls.clear()
    d = new("DelayBlock", 129, C)
    meth(d, "initialize")
    assert lib.calls == [("lrb200_delay_create", (129, 8, 0))] and meth(d, "gpu_capable") == [True]
    meth(d, "process", x)
    assert [c[0] for c in lib.calls][1:] == ["lrb200_block_max_output", "lrb200_block_execute"]
    lib.calls.clear()
    d = new("DelayBlock", 5, F)
    meth(d, "make_device_handle")
    assert lib.calls == [("lrb200_delay_create", (5, 4, 1))]
    lib.calls.clear()
    d = new("DelayBlock", 7, B)
    meth(d, "initialize")
    meth(d, "process", x)
    assert lib.calls == [] and d.hash["host_initialized"] is True and d.hash["host_processed"] == 1
    assert meth(d, "gpu_capable") == [False]
    # PLLBlock: Hz arguments and the block's rate; two outputs
    lib.calls.clear()
    p = new("PLLBlock", 100.0, 18950.0, 19050.0, 2.0)
    p.hash["rate"] = 220500.0
    meth(p, "initialize")
    assert lib.calls == [("lrb200_pll_create", (100.0, 18950.0, 19050.0, 2.0, 220500.0, 0))]
    lib.calls.clear()
    out = meth(p, "process", x)
    assert len(out) == 2 and out[0].hash["data_type"] is C and out[1].hash["data_type"] is F
    assert lib.calls[1][0] == "lrb200_block_execute_multi" and lib.calls[1][1][2:6:3] == (1, 2)
    p = new("PLLBlock", 100.0, 18950.0, 19050.0)
    p.hash["rate"] = 1.0
    meth(p, "make_device_handle")
    assert lib.calls[-1][1][3] == 1.0 and lib.calls[-1][1][5] == 1            # default multiplier; DEVICE pointers


def dag_topologies():
    import luaradio_b200 as radio
    x = np.zeros(16, np.complex64)
    tops = dict(topologies())
    top = radio.CompositeBlock()
    top.connect(radio.ArraySource(x, 1e6), radio.AMSynchronousDemodulator(100e3, 5e3), radio.ArraySink())
    top._prepare_to_run(initialize=False)
    tops["am_synchronous"] = top
    return tops


@pytest.mark.parametrize("name", ["stereo", "am_synchronous", "mono", "two_sources", "host_and_fanout"])
def test_lua_dag_planner_matches_the_python_planner(monkeypatch, name):
    """plan_gpu_dags / collapse_gpu_dags (Lua, executed) against CompositeBlock._plan_gpu_dags (Python): same member sets,
    same outside feed, same outside-read outputs; the rewritten map is consistent; no linear run is left among the members."""
    it, lib, types, radio = patched_radio(monkeypatch)
    top = dag_topologies()[name]
    lua_gpu = {b: LUA_GPU_BASE[b.name] for b in top._concrete_order if b.name in LUA_GPU_BASE}
    from luaradio_b200.signal_blocks import GPUBlock
    assert all(b in lua_gpu for b in top._concrete_order if isinstance(b, GPUBlock) and b.inputs and b.outputs), "a Python GPU block without a Lua form"
    expected = top._plan_gpu_dags()
    lua_of, conns = export_graph(it, radio, types, top, lua_gpu)
    name_of = {id(lb): b for b, lb in lua_of.items()}
    patch = it.require("radio_b200.composite_patch")
    plans = it.call(patch.hash["plan_gpu_dags"], [conns])[0].array()
    assert len(plans) == len(expected)
    assert (expected == []) == (name in ("mono", "two_sources"))          # host_and_fanout: Downsampler -> {ComplexMagnitude, ComplexToReal}
    if not expected:
        return
    (members_py, ext_in_py, ext_out_py), plan = expected[0], plans[0]
    members = plan.hash["members"].array()
    assert {id(m) for m in members} == {id(lua_of[b]) for b in members_py} and len(members) == len(members_py)
    assert name_of[id(plan.hash["ext_in"].hash["owner"])] is ext_in_py.owner
    assert sorted(name_of[id(p.hash["owner"])].name for p in plan.hash["ext_out"].array()) == sorted(p.owner.name for p in ext_out_py)
    # evaluation order: every member comes after the members that feed it
    index = {id(m): k for k, m in enumerate(members)}
    for m in members:
        for p in m.hash["inputs"].array():
            up = conns.hash[p].hash["owner"]
            if id(up) in index:
                assert index[id(up)] < index[id(m)]
    # rewrite
    before = dict(conns.hash)
    it.call(patch.hash["collapse_gpu_dags"], [conns])
    dags = {id(o.hash["owner"]): o.hash["owner"] for o in conns.hash.values() if "ext_out" in o.hash["owner"].hash}
    assert len(dags) == 1
    dag = next(iter(dags.values()))
    member_ids = {id(m) for m in members}
    for i, o in conns.hash.items():
        assert id(i.hash["owner"]) not in member_ids and id(o.hash["owner"]) not in member_ids
    assert conns.hash[dag.hash["inputs"].hash[1]] is plan.hash["ext_in"]
    ext_out = plan.hash["ext_out"].array()
    for i, o in before.items():
        if id(i.hash["owner"]) in member_ids:
            continue
        if id(o.hash["owner"]) in member_ids:
            k = [id(p) for p in ext_out].index(id(o))
            assert conns.hash[i] is dag.hash["outputs"].hash[k + 1]
            assert dag.hash["outputs"].hash[k + 1].hash["data_type"] is o.hash["data_type"]
        else:
            assert conns.hash[i] is o
    # the chain planner then finds nothing among what is left (stereo / AM synchronous: source -> dag -> sinks)
    n_before = len(conns.hash)
    it.call(patch.hash["collapse_gpu_runs"], [conns])
    chains = {id(o.hash["owner"]): o.hash["owner"] for o in conns.hash.values() if "blocks" in o.hash["owner"].hash and "ext_out" not in o.hash["owner"].hash}
    if name == "host_and_fanout":           # the run in front of the host block is still a chain of its own
        assert [[b.hash["name"] for b in c.hash["blocks"].array()] for c in chains.values()] == [["FrequencyTranslatorBlock", "LowpassFilterBlock"]]
    else:
        assert len(conns.hash) == n_before and not chains
    py_runs = [[b.name for b in run] for run, _, _ in top._plan_gpu_runs({m for m in members_py})]
    assert py_runs == ([["FrequencyTranslatorBlock", "LowpassFilterBlock"]] if name == "host_and_fanout" else [])


@pytest.mark.parametrize("name", ["stereo", "am_synchronous"])
def test_dag_block_builds_a_valid_device_dag(monkeypatch, name):
    """GPUDagBlock:initialize() against the mock library: dag_create, then nodes whose input references are -1 (the DAG's
    input) or node * 4 + k of an EARLIER node with that many outputs; every member lands in exactly one node; linear runs
    inside the set are committed as fused flow graphs; set_outputs names valid references; process() returns one vector per
    outside-read output."""
    it, lib, types, radio = patched_radio(monkeypatch)
    top = dag_topologies()[name]
    lua_gpu = {b: LUA_GPU_BASE[b.name] for b in top._concrete_order if b.name in LUA_GPU_BASE}
    lua_of, conns = export_graph(it, radio, types, top, lua_gpu)
    patch = it.require("radio_b200.composite_patch")
    it.call(patch.hash["collapse_gpu_dags"], [conns])
    dag = next(o.hash["owner"] for o in conns.hash.values() if "ext_out" in o.hash["owner"].hash)
    meth = lambda obj, nm, *a: it.call(it.index(obj, nm), [obj] + list(a))
    lib.calls.clear()
    lib.dag_nodes = 0
    meth(dag, "initialize")
    assert lib.calls[0][0] == "lrb200_dag_create"
    d = dag.hash["dag"]
    n_members = dag.hash["blocks"].length()
    node_outputs, placed, pending_graph = [], 0, {}
    for nm, args in lib.calls[1:]:
        if nm == "lrb200_graph_append":
            pending_graph[id(args[0])] = pending_graph.get(id(args[0]), 0) + 1
            assert args[1].args[-1] == 1                               # DEVICE pointers
        elif nm == "lrb200_graph_commit":
            assert args[1] == 1 and pending_graph[id(args[0])] >= 2
        elif nm == "lrb200_dag_add_graph":
            assert args[0] is d
            ref = args[2]
            assert ref == -1 or (ref // 4 < len(node_outputs) and ref % 4 < node_outputs[ref // 4])
            node_outputs.append(1)
            placed += pending_graph.pop(id(args[1]))
        elif nm == "lrb200_dag_add_block":
            assert args[0] is d and args[1].args[-1] == 1
            for k in range(args[3]):
                ref = args[2].hash[k]
                assert ref == -1 or (ref // 4 < len(node_outputs) and ref % 4 < node_outputs[ref // 4]), (nm, ref)
            node_outputs.append(2 if args[1].what == "lrb200_pll_create" else 1)
            assert args[3] == (2 if args[1].what == "lrb200_binary_create" else 1)
            placed += 1
    assert placed == n_members and not pending_graph
    assert any(c[0] == "lrb200_dag_add_graph" for c in lib.calls)      # the linear runs inside the set are fused graphs
    so = [c for c in lib.calls if c[0] == "lrb200_dag_set_outputs"]
    n_ext = dag.hash["ext_out"].length()
    assert len(so) == 1 and so[0][1][2] == n_ext
    for k in range(n_ext):
        ref = so[0][1][1].hash[k]
        assert ref >= 0 and ref // 4 < len(node_outputs) and ref % 4 < node_outputs[ref // 4]
    assert lib.calls[-1][0] == "lrb200_dag_set_outputs"
    # process
    lib.calls.clear()
    x = vec(types, "ComplexFloat32", 4096)
    outs = meth(dag, "process", x)
    assert len(outs) == n_ext and all(o.hash["length"] == 4096 for o in outs)
    assert [c[0] for c in lib.calls] == ["lrb200_dag_max_output"] * n_ext + ["lrb200_dag_execute"]
    assert lib.calls[-1][1][:3] == (d, x.hash["data"], 4096)
    assert meth(dag, "get_rate") == [1e6]
