#!/usr/bin/env python3
"""Extract the reference's committed golden vectors into portable fixtures.

Run HERE (build container, where /root/reference exists):

    python tests/golden/make_golden.py

It parses the reference's generated busted specs (`tests/**/*_spec.gen.lua`,
written by the reference's `tests/generate.py:82-126`) and the utilities
vector modules (`tests/utilities/*_vectors.gen.lua`, `tests/top_vectors.gen.lua`)
and writes one `.npz` per spec under `tests/golden/`.  Nothing is computed:
these are the reference's own numbers, byte for byte (8-decimal text -> float32).

Each BlockSpec npz holds
    manifest : JSON string  {"block", "epsilon", "vectors": [{"desc", "args", "n_in", "n_out"}]}
               (array-valued args are replaced by {"array": "v<i>_arg<k>"})
    v<i>_in<j>, v<i>_out<j>, v<i>_arg<k> : numpy arrays (float32 / complex64)

/root/reference does not exist on the GPU box, so the fixtures are committed.
"""
import json
import os
import re
import sys

import numpy as np

REF = os.environ.get("LUARADIO_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

BLOCK_SPECS = [
    "blocks/signal/firfilter_spec", "blocks/signal/lowpassfilter_spec", "blocks/signal/highpassfilter_spec",
    "blocks/signal/bandpassfilter_spec", "blocks/signal/bandstopfilter_spec",
    "blocks/signal/complexbandpassfilter_spec", "blocks/signal/complexbandstopfilter_spec",
    "blocks/signal/hilberttransform_spec", "blocks/signal/frequencytranslator_spec",
    "blocks/signal/frequencydiscriminator_spec", "blocks/signal/downsampler_spec",
    "blocks/signal/fmdeemphasisfilter_spec", "blocks/signal/singlepolelowpassfilter_spec",
    "blocks/signal/singlepolehighpassfilter_spec", "blocks/signal/iirfilter_spec",
    "blocks/signal/complexmagnitude_spec", "blocks/signal/complextoreal_spec",
    "blocks/signal/multiplyconjugate_spec",
    "composites/tuner_spec", "composites/decimator_spec",
    "blocks/signal/multiplyconstant_spec", "blocks/signal/upsampler_spec",
    "composites/interpolator_spec", "composites/rationalresampler_spec",
    "blocks/signal/multiply_spec", "blocks/signal/add_spec", "blocks/signal/subtract_spec", "blocks/signal/delay_spec",
]
MODULE_SPECS = ["utilities/filter_utils_vectors", "utilities/window_utils_vectors", "utilities/spectrum_utils_vectors"]


class LuaLit:
    """Tiny recursive-descent parser for the literal subset generate.py emits."""

    def __init__(self, s, pos=0):
        self.s, self.p = s, pos

    def ws(self):
        while self.p < len(self.s) and self.s[self.p] in " \t\r\n":
            self.p += 1

    def peek(self, tok):
        self.ws()
        return self.s.startswith(tok, self.p)

    def eat(self, tok):
        self.ws()
        if not self.s.startswith(tok, self.p):
            raise ValueError("expected %r at %d: %r" % (tok, self.p, self.s[self.p:self.p + 40]))
        self.p += len(tok)

    def value(self):
        self.ws()
        s = self.s
        for prefix, dtype in (("radio.types.ComplexFloat32.vector_from_array(", np.complex64),
                              ("radio.types.Float32.vector_from_array(", np.float32),
                              ("radio.types.Bit.vector_from_array(", np.uint8),
                              ("radio.types.Byte.vector_from_array(", np.uint8)):
            if s.startswith(prefix, self.p):
                self.p += len(prefix)
                arr = self.value()
                self.eat(")")
                if dtype is np.complex64:
                    a = np.array([complex(np.float32(r), np.float32(i)) for r, i in arr], dtype=np.complex64)
                else:
                    a = np.array(arr, dtype=dtype)
                return a.reshape(-1)
        for prefix, is_c in (("radio.types.ComplexFloat32(", True), ("radio.types.Float32(", False)):
            if s.startswith(prefix, self.p):
                self.p += len(prefix)
                vals = [self.value()]
                while self.peek(","):
                    self.eat(",")
                    vals.append(self.value())
                self.eat(")")
                return {"complex": vals} if is_c else float(vals[0])
        if s[self.p] == "{":
            self.p += 1
            items, keyed = [], {}
            while True:
                self.ws()
                if s[self.p] == "}":
                    self.p += 1
                    break
                m = re.compile(r"([A-Za-z_][A-Za-z_0-9]*)\s*=\s*").match(s, self.p)
                if m:
                    self.p = m.end()
                    keyed[m.group(1)] = self.value()
                else:
                    items.append(self.value())
                self.ws()
                if s[self.p] == ",":
                    self.p += 1
            return keyed if keyed else items
        if s[self.p] == '"':
            e = s.index('"', self.p + 1)
            v = s[self.p + 1:e]
            self.p = e + 1
            return v
        m = re.compile(r"true|false|nil").match(s, self.p)
        if m:
            self.p = m.end()
            return {"true": True, "false": False, "nil": None}[m.group(0)]
        m = re.compile(r"0x[0-9a-fA-F]+|[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?").match(s, self.p)
        if m:
            self.p = m.end()
            t = m.group(0)
            if t.startswith("0x"):
                return int(t, 16)
            return float(t) if re.search(r"[.eE]", t) else int(t)
        raise ValueError("cannot parse at %d: %r" % (self.p, s[self.p:self.p + 60]))


def parse_block_spec(path):
    s = open(path).read()
    m = re.search(r"jigs\.TestBlock\(radio\.(\w+), ", s)
    block = m.group(1)
    p = LuaLit(s, m.end())
    vectors = p.value()
    p.eat(",")
    p.eat("{")
    p.eat("epsilon")
    p.eat("=")
    e = s.index("}", p.p)
    epsilon = s[p.p:e].strip()
    return block, vectors, epsilon


def jsonable_arg(a, arrays, key):
    if isinstance(a, np.ndarray):
        arrays[key] = a
        return {"array": key}
    if isinstance(a, list):
        return [jsonable_arg(x, arrays, key + "_%d" % i) for i, x in enumerate(a)]
    return a


def main():
    n_vec = 0
    for spec in BLOCK_SPECS:
        block, vectors, epsilon = parse_block_spec(os.path.join(REF, "tests", spec + ".gen.lua"))
        arrays, man = {}, {"block": block, "epsilon": epsilon, "source": "tests/" + spec + ".gen.lua", "vectors": []}
        for i, v in enumerate(vectors):
            args = [jsonable_arg(a, arrays, "v%d_arg%d" % (i, k)) for k, a in enumerate(v["args"])]
            for j, a in enumerate(v["inputs"]):
                arrays["v%d_in%d" % (i, j)] = a
            for j, a in enumerate(v["outputs"]):
                arrays["v%d_out%d" % (i, j)] = a
            man["vectors"].append({"desc": v["desc"], "args": args, "n_in": len(v["inputs"]), "n_out": len(v["outputs"])})
            n_vec += 1
        arrays["manifest"] = np.array(json.dumps(man))
        np.savez_compressed(os.path.join(OUT, os.path.basename(spec) + ".npz"), **arrays)
        print("%-34s %-30s %3d vectors  eps=%s" % (os.path.basename(spec), block, len(vectors), epsilon))

    for spec in MODULE_SPECS:
        s = open(os.path.join(REF, "tests", spec + ".gen.lua")).read()
        arrays = {}
        for m in re.finditer(r"^M\.(\w+) = ", s, re.M):
            arrays[m.group(1)] = LuaLit(s, m.end()).value()
        np.savez_compressed(os.path.join(OUT, os.path.basename(spec) + ".npz"), **arrays)
        print("%-34s %d arrays" % (os.path.basename(spec), len(arrays)))

    # tests/blocks/sources/iqfile_spec.gen.lua: the args hold the raw file bytes as a Lua "\x.." string
    # (iqfile_spec.py:47-52); keep bytes + format + expected complex64 output per vector
    s = open(os.path.join(REF, "tests", "blocks/sources/iqfile_spec.gen.lua")).read()
    arrays, fmts = {}, []
    for i, m in enumerate(re.finditer(r'args = \{require\(\'tests\.buffer\'\)\.open\("([^"]*)"\), "(\w+)", 1\},\s*inputs = \{\},\s*outputs = \{', s)):
        raw = bytes(int(h, 16) for h in re.findall(r"\\x([0-9a-f]{2})", m.group(1)))
        arrays["v%d_raw" % i] = np.frombuffer(raw, dtype=np.uint8).copy()
        arrays["v%d_out" % i] = LuaLit(s, m.end()).value()
        fmts.append(m.group(2))
    arrays["formats"] = np.array(fmts)
    np.savez_compressed(os.path.join(OUT, "iqfile_spec_raw.npz"), **arrays)
    print("iqfile_spec_raw: %d formats: %s" % (len(fmts), " ".join(fmts)))

    # tests/blocks/sources/realfile_spec.gen.lua: same layout, one component per sample, Float32 outputs
    s = open(os.path.join(REF, "tests", "blocks/sources/realfile_spec.gen.lua")).read()
    arrays, fmts = {}, []
    for i, m in enumerate(re.finditer(r'args = \{require\(\'tests\.buffer\'\)\.open\("([^"]*)"\), "(\w+)", 1\},\s*inputs = \{\},\s*outputs = \{', s)):
        raw = bytes(int(h, 16) for h in re.findall(r"\\x([0-9a-f]{2})", m.group(1)))
        arrays["v%d_raw" % i] = np.frombuffer(raw, dtype=np.uint8).copy()
        arrays["v%d_out" % i] = LuaLit(s, m.end()).value()
        fmts.append(m.group(2))
    arrays["formats"] = np.array(fmts)
    np.savez_compressed(os.path.join(OUT, "realfile_spec_raw.npz"), **arrays)
    print("realfile_spec_raw: %d formats: %s" % (len(fmts), " ".join(fmts)))

    # tests/top_vectors.gen.lua: raw little-endian byte strings ("\x.."), tests/top_vectors.py:26-36
    s = open(os.path.join(REF, "tests", "top_vectors.gen.lua")).read()
    arrays = {}
    for name, dtype in (("SRC1_TEST_VECTOR", np.complex64), ("SRC2_TEST_VECTOR", np.complex64), ("SNK_TEST_VECTOR", np.float32)):
        m = re.search(r'M\.%s = "([^"]*)"' % name, s)
        raw = bytes(int(h, 16) for h in re.findall(r"\\x([0-9a-f]{2})", m.group(1)))
        arrays[name] = np.frombuffer(raw, dtype=dtype).copy()
    np.savez_compressed(os.path.join(OUT, "top_vectors.npz"), **arrays)
    print("top_vectors: src %d, snk %d" % (arrays["SRC1_TEST_VECTOR"].size, arrays["SNK_TEST_VECTOR"].size))
    print("total block vectors:", n_vec)


if __name__ == "__main__":
    sys.exit(main())
