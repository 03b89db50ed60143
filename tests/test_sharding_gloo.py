"""CPU tier: the N > 1 path with world_size 2 over gloo. Each rank converts its shard with the host emulation
of the engine (tests/emu): streams sharded by stream (BASELINE config 4 shape) and one stream time-chunked
with halos (config 5 shape); results are gathered with the same collective wrapper the GPU path uses (NCCL
there, gloo here) and compared with the unsharded conversion, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from foo_dsp_resampler_b200 import sharding  # noqa: E402


def test_partition_helpers():
    assert [sharding.stream_shard(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sum(c for _, c in (sharding.stream_shard(4096, 8, r) for r in range(8))) == 4096
    chunks = sharding.time_chunks(100000, 8, 1766)
    assert chunks[0][0] == 0 and sum(c for _, c in chunks) == 100000
    assert all(b % 1766 == 0 for b, _ in chunks)
    assert all(chunks[r][0] + chunks[r][1] == chunks[r + 1][0] for r in range(7))
    assert sharding.time_chunks(10, 4, 1000) == [(0, 0), (0, 0), (0, 0), (0, 10)] or \
        sum(c for _, c in sharding.time_chunks(10, 4, 1000)) == 10


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emulib
    import signals
    from foo_dsp_resampler_b200 import _capi, converter
    L = emulib.lib()
    ok = True
    # --- streams sharded by stream ---
    cfg = _capi.make_config(48000, 44100)
    nstreams, nch, n = 6, 2, 9000
    xs = np.stack([signals.sweep_noise(48000, nch, n, stream=s) for s in range(nstreams)])
    first, count = sharding.stream_shard(nstreams, world, rank)
    b = converter.BatchConverter(cfg, nch, count, n, engine="float", lib=L)
    nout = b.frames_out(n)
    mine = np.zeros((count, nout, nch), np.float32)
    b.process(np.ascontiguousarray(xs[first:first + count]).ctypes.data, n, mine.ctypes.data)
    b.close()
    parts = sharding.gather_outputs(torch.from_numpy(mine))
    if rank == 0:
        full = converter.BatchConverter(cfg, nch, nstreams, n, engine="float", lib=L)
        ref = np.zeros((nstreams, nout, nch), np.float32)
        full.process(xs.ctypes.data, n, ref.ctypes.data)
        full.close()
        ok &= bool(np.array_equal(np.concatenate([p.numpy() for p in parts]), ref))
    # --- one stream, time-chunked with halos (384 kHz -> 48 kHz: h12, h12, DFT /2) ---
    cfg = _capi.make_config(384000, 48000)
    nch, n = 2, 160000
    x = signals.sweep_noise(384000, nch, n)[None]
    b = converter.BatchConverter(cfg, nch, 1, n, engine="float", lib=L)
    nout = b.frames_out(n)
    chunks = sharding.time_chunks(nout, world, sharding.last_stage_block(b.plan()))
    ob, oc = chunks[rank]
    f, c = b.input_window(n, ob, oc)
    assert c < n                                     # each rank reads only its halo'd window
    win = np.ascontiguousarray(x[:, f:f + c, :])
    part = np.zeros((1, oc, nch), np.float32)
    b.process_range(win.ctypes.data, f, c, n, ob, oc, part.ctypes.data)
    width = max(cc for _, cc in chunks)
    padded = np.zeros((1, width, nch), np.float32)
    padded[:, :oc] = part
    parts = sharding.gather_outputs(torch.from_numpy(padded))
    if rank == 0:
        ref = np.zeros((1, nout, nch), np.float32)
        b.process(x.ctypes.data, n, ref.ctypes.data)
        got = np.concatenate([p.numpy()[:, :chunks[r][1]] for r, p in enumerate(parts)], axis=1)
        ok &= bool(np.array_equal(got, ref))
    b.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ret[rank] = int(flag.item())
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    import emulib
    emulib.lib()                                     # build the emulation library once, before forking
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert dict(ret) == {0: 1, 1: 1}
