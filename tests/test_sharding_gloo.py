"""World-size-2 gloo test (CPU) of the time-chunk sharding host logic: partition, halo size, the P2P halo
exchange and output trimming reproduce the single-stream result.  The per-chunk arithmetic here is the
ORACLE chain (test infrastructure) standing in for the GPU graph; the GPU version of the same flow is
bench.py --gpus N."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luaradio_b200 import sharding
from oracle import lr_oracle as O

N_TOTAL = 250000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = sharding.plan_chunks(N_TOTAL, world, 25)
    start, count = plan[rank]
    b, a = O.fm_deemphasis_taps(75e-6, 220500.0)
    halo = sharding.chain_halo([(128, 1), (2, 5), (128, 5)], iir_pole=float(-a[1]), iir_rate_div=5)
    x_local = torch.view_as_real(torch.from_numpy(O.synth_fm_iq(start, count)))      # gloo has no complex dtype
    halo_buf = torch.zeros(halo, 2)
    sharding.exchange_halo(dist, x_local, halo_buf, rank, world, halo)
    lead = halo if rank > 0 else 0
    xin = torch.view_as_complex(torch.cat([halo_buf[halo - lead:], x_local]).contiguous()).numpy()
    chain = O.wbfm_mono_chain()
    chain.blocks[0].blocks[0].n0 = start - lead                  # translator phase from the global sample index
    y = chain.process(xin)
    skip, keep = sharding.trim_outputs(len(y), lead, 25)
    np.save(os.path.join(out_dir, "y%d.npy" % rank), y[skip:])
    dist.barrier()
    dist.destroy_process_group()


def test_plan_and_halo():
    plan = sharding.plan_chunks(1000003, 4, 25)
    assert plan[0][0] == 0 and sum(c for _, c in plan) == 1000003
    assert all(s % 25 == 0 for s, _ in plan)
    assert all(plan[i][0] + plan[i][1] == plan[i + 1][0] for i in range(3))
    h = sharding.chain_halo([(128, 1), (2, 5), (128, 5)], iir_pole=0.9413, iir_rate_div=5)
    assert h % 25 == 0 and 3000 < h <= 4000          # bench.py's HALO = 4000 covers it
    with pytest.raises(ValueError):
        sharding.plan_chunks(10, 4, 25)


def test_two_rank_halo_exchange_reproduces_single_stream(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    y = np.concatenate([np.load(tmp_path / "y0.npy"), np.load(tmp_path / "y1.npy")])
    ref = O.wbfm_mono_chain().process(O.synth_fm_iq(0, N_TOTAL))
    assert y.shape == ref.shape
    assert np.max(np.abs(y - ref)) <= 1e-5 * max(1.0, np.max(np.abs(ref)))
