"""CPU, world_size 2 over gloo: the multi-GPU plumbing (frame i -> rank i mod R, one gather of the decoded frames).
The decode itself is the CPU emulation of the kernels here; on GPUs the same functions run over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _emul_decode(files, num_channels, dtype):
    import emul_lib
    import jxlo
    dt = {np.dtype(np.uint8): jxlo.UINT8, np.dtype(np.uint16): jxlo.UINT16}[np.dtype(dtype)]
    shapes = []
    for f in files:
        d = jxlo.Decoded(f)
        shapes.append((d.info.ysize, d.info.xsize))
    return emul_lib.decode(list(files), num_channels, dt, shapes)


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    import jxlo
    import vardct_cases as vc
    from conftest import read_golden
    pkg = ge.load_package()
    files = [vc.encoded("dct8_filters")[0], read_golden("sample.jxl"), vc.encoded("odd_size")[0],
             vc.encoded("heuristic")[0], read_golden("sample.jxl")]
    assert pkg.shard_indices(5, rank, world) == list(range(rank, 5, world))
    outs = pkg.decode_batch_distributed(files, 4, np.uint8, dst=0, decode_fn=_emul_decode)
    if rank == 0:
        ok = all(np.array_equal(o, jxlo.decode(f, 4, jxlo.UINT8)) for o, f in zip(outs, files))
        q.put(bool(ok) and len(outs) == 5)
    else:
        assert outs is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_shard_indices_cover_every_frame_once():
    import __graft_entry__ as ge
    pkg = ge.load_package()
    for n in (0, 1, 7, 512):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in pkg.shard_indices(n, r, world))
            assert got == list(range(n))
