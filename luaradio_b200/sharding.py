"""Time-chunk sharding of a sample stream across GPUs.

The hot path has no cross-sample coupling beyond a finite memory: FIR histories, one discriminator
sample and an exponentially decaying IIR state.  So the stream is cut into contiguous chunks on
multiples of `align` input samples (the product of the decimation factors, so every decimator keeps
its phase), rank r > 0 additionally receives the last `halo` input samples of rank r-1's chunk from its
left neighbour (one point-to-point message per step: NCCL send/recv over NVLink on GPUs, gloo in the
CPU tests), runs from a cold state `halo` samples early and drops the outputs that belong to the halo.
There is no other exchange on this path; outputs are disjoint slices.

This module is transport-agnostic host logic (torch.distributed tensors, any backend); the kernels never
see it.  bench.py uses it for the N > 1 runs.
"""
import math


def plan_chunks(total, world, align=25):
    """Contiguous [start, start+count) per rank; every boundary is a multiple of `align`."""
    per = (total // world) // align * align
    if per <= 0:
        raise ValueError("stream too short to shard: total=%d world=%d align=%d" % (total, world, align))
    plan = []
    for r in range(world):
        start = r * per
        count = per if r < world - 1 else total - start
        plan.append((start, count))
    return plan


def chain_halo(fir_taps_by_rate, iir_pole=None, iir_rate_div=1, tol=1e-12, align=25):
    """Input samples of lead-in needed so a cold start is indistinguishable (to `tol`) from the stream.

    fir_taps_by_rate: [(ntaps, rate_divisor)]: a FIR with ntaps at input_rate / rate_divisor needs
    (ntaps - 1) * rate_divisor input samples; a discriminator needs 1 sample at its rate (pass ntaps=2).
    iir_pole: |c| of a single-pole recurrence running at input_rate / iir_rate_div.
    """
    need = 0
    for ntaps, div in fir_taps_by_rate:
        need += (ntaps - 1) * div
    if iir_pole:
        warm = int(math.ceil(math.log(tol) / math.log(abs(iir_pole)))) if abs(iir_pole) < 1 else 0
        need += warm * iir_rate_div
    return int(math.ceil(need / align) * align)


def exchange_halo(dist, chunk, halo_buf, rank, world, halo):
    """Rank r sends the last `halo` samples of `chunk` to r+1 and receives its own halo from r-1 into
    `halo_buf` (both 1-D tensors on the backend's device).  One batched P2P op per neighbour."""
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, chunk[chunk.shape[0] - halo:], rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, halo_buf, rank - 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def trim_outputs(n_out_total, lead, total_decimation):
    """Outputs produced from `lead` halo inputs (lead is a multiple of the total decimation) to drop."""
    assert lead % total_decimation == 0
    skip = lead // total_decimation
    return skip, n_out_total - skip
