"""Multi-GPU partitioning of the rate path. No data-path collective exists: ranks own
disjoint units and the only communication is the optional gather of results.

* batched workloads: streams are independent (they share only read-only coefficient banks,
  /root/reference/rate/rate_base.h:89-92) -> ``stream_shard`` gives every rank a contiguous slice;
* one long stream: every stage is FIR and every block / phase position is a closed-form function of the absolute
  sample index, so the OUTPUT timeline is cut into contiguous ranges (``time_chunks``); each rank asks the
  engine which input window its range depends on (filter-history halo, RRX_batch_input_window) and the engine
  starts it at the exactly computed phase (RRX_batch_process_range). Results are bit-identical to the
  single-device run."""


def stream_shard(nstreams, world, rank):
    """Contiguous slice (first, count) of ``nstreams`` for ``rank``; the remainder goes to the low ranks."""
    base, extra = divmod(int(nstreams), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def time_chunks(nout_total, world, align=1):
    """Cut output frames [0, nout_total) into ``world`` contiguous ranges whose interior boundaries are
    multiples of ``align`` (e.g. the output block size of the last DFT stage, so no block is split between two
    ranks and computed twice). Returns [(out_begin, out_count)] per rank; trailing ranks may be empty."""
    nout_total, world, align = int(nout_total), int(world), max(1, int(align))
    units = -(-nout_total // align)
    bounds = [min(nout_total, ((units * r) // world) * align) for r in range(world)] + [nout_total]
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def last_stage_block(plan):
    """Output frames produced per work unit of the last stage (alignment hint for time_chunks)."""
    st = plan["stages"][-1] if plan["stages"] else None
    if not st or st["kind"] != 1:
        return 1
    valid = st["dft_length"] - (st["num_taps"] - 1)
    step = st["step_int"]
    if step == 1:
        return valid
    if step > 1:
        return 1
    return valid >> (-step)


def gather_outputs(local, world_group=None):
    """All-gather equally shaped per-rank result tensors (NCCL on GPUs, gloo on CPU). The one collective on
    the path: the caller asks for it only if the results are needed on every / one device."""
    import torch
    import torch.distributed as dist
    bufs = [torch.empty_like(local) for _ in range(dist.get_world_size(world_group))]
    dist.all_gather(bufs, local, group=world_group)
    return bufs
