"""Build PRODUCT blocks (luaradio_b200) for golden vectors the way the reference's jig does
(tests/jigs.lua:60-84): instantiate(args), get_rate -> 2.0, differentiate(input types), initialize."""
import numpy as np

import luaradio_b200 as radio
from luaradio_b200.types import ComplexFloat32, Float32, Vector


def _lua_arg(a):
    # golden args: numpy arrays are Vector args (taps); dicts are option tables
    if isinstance(a, np.ndarray):
        return (ComplexFloat32 if np.iscomplexobj(a) else Float32).vector_from_array(a)
    if isinstance(a, dict) and "complex" in a:          # radio.types.ComplexFloat32(re, im) constant
        return complex(a["complex"][0], a["complex"][1])
    return a


def create_block(block_name, args, inputs, rate=2.0):
    cls = getattr(radio, block_name)
    blk = cls(*[_lua_arg(a) for a in args])
    blk.get_rate = lambda: rate
    blk.differentiate([ComplexFloat32 if np.iscomplexobj(x) else Float32 for x in inputs])
    blk.initialize()
    return blk


def run_whole(blk, x):
    out = blk.process(Vector.cast(x))
    return np.array(out.data, copy=True)


def run_sample_by_sample(blk, x, out_dtype):
    outs = []
    for i in range(len(x)):
        o = blk.process(Vector.cast(x[i:i + 1]))
        outs.append(np.array(o.data, copy=True))
    return np.concatenate(outs) if outs else np.zeros(0, out_dtype)


def run_composite(block_name, args, x, rate=2.0, fuse=True, chunk=1 << 22):
    """tests/jigs.lua:89-147: source -> composite -> sink as a real flow graph, single process."""
    src = radio.ArraySource(x, rate, chunk)
    snk = radio.ArraySink()
    blk = getattr(radio, block_name)(*[_lua_arg(a) for a in args])
    top = radio.CompositeBlock()
    top.connect(src, blk, snk)
    top.run(False, fuse=fuse)
    return snk.result(), top
