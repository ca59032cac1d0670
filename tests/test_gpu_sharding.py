"""Time-chunk sharding on the GPU path (single device standing in for N ranks): each 'rank' runs the fused chain
from a cold state HALO samples before its chunk (lrb200_graph_seek), drops the halo outputs, and the
concatenation must equal the single-stream result -- the property bench.py --gpus N relies on."""
import ctypes
import math

import numpy as np
import pytest

from luaradio_b200 import _lib, sharding
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu


def build_graph(lib):
    import bench
    return bench.build_chain_graph(lib, _lib)


def run_graph(lib, g, x, seek):
    _lib.check(lib.lrb200_graph_reset(g))
    _lib.check(lib.lrb200_graph_seek(g, seek))
    out = np.zeros(lib.lrb200_graph_max_output(g, len(x)) + 8, np.float32)
    no = ctypes.c_size_t()
    _lib.check(lib.lrb200_graph_execute(g, x.ctypes.data, len(x), out.ctypes.data, ctypes.byref(no)))
    return out[:no.value]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_chain_equals_single_stream(world):
    import bench
    lib = _lib.require_device()
    total = 1500000
    x = O.synth_fm_iq(0, total)
    g = build_graph(lib)
    whole = run_graph(lib, g, x, 0)
    ref = O.wbfm_mono_chain().process(x)
    assert np.max(np.abs(whole - ref)) <= 1e-5
    parts = []
    for start, count in sharding.plan_chunks(total, world, 25):
        lead = bench.HALO if start > 0 else 0
        y = run_graph(lib, g, np.ascontiguousarray(x[start - lead:start + count]), start - lead)
        skip, keep = sharding.trim_outputs(len(y), lead, 25)
        parts.append(y[skip:])
    got = np.concatenate(parts)
    assert got.shape == whole.shape
    # halo restart vs carried state: identical up to float32 rounding of the scan order
    assert np.max(np.abs(got - whole)) <= 2e-6
    lib.lrb200_graph_destroy(g)
