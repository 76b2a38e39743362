"""One process per rank, as in production (bench.py N > 1): the sharded ParseND whose exchange is the library's kernel over
peer memory, with the ranks' buffers opened through CUDA IPC (sj_exchange_create / sj_exchange_connect via
parallel.ShardedParse.connect_exchange; the handles travel over gloo).  Two processes; they share GPU 0 when the box has
only one.  The slices laid end to end must be the oracle's tape and string buffer of the whole stream."""
import os
import sys

import numpy as np
import pytest

from tests.test_multi_rank_cpu import _free_port
from tests.util import load_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _stream():
    pk = load_fixture("parking-citations").strip()
    return b"\n".join([pk] * 3) + b'\n{"esc":"a\\u00e9\\n","n":[1,2.5,-3],"t":true}\n' + pk[:40000].rsplit(b"\n", 1)[0]


def _worker(rank, world, port, copy, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "simdjson-go_b200"))
    import torch
    import torch.distributed as dist
    import simdjson_b200 as sj
    from simdjson_b200.parallel import ShardedParse, split_at_newlines, trimmed_window
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import datetime
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    try:
        devno = rank % torch.cuda.device_count()
        torch.cuda.set_device(devno)
        dev = torch.device("cuda", devno)
        ctx = sj.Context(devno)
        stream = _stream()
        wins = [trimmed_window(stream, a, b) for a, b in split_at_newlines(stream, world)]
        a, b = wins[rank]
        sp = ShardedParse(ctx)  # (CPU tensors over gloo for the one-time exchange of the handles)
        rc = sp.connect_exchange(rank, world, gap_bytes=wins[rank + 1][0] - b if rank + 1 < world else 0)
        if rc != 0:
            open(os.path.join(outdir, "unsupported%d" % rank), "w").write(str(rc))
            return
        assert ctx.L.sj_exchange_set_timeout_ms(ctx.h, 30000) == 0
        d_msg = torch.full((b - a + 256,), 0x20, dtype=torch.uint8, device=dev)
        d_msg[: b - a] = torch.frombuffer(bytearray(stream[a:b]), dtype=torch.uint8).to(dev)
        torch.cuda.synchronize()
        for it in range(3):  # (epochs: the slots are double-buffered)
            dist.barrier()
            rc, tot = sp.count(d_msg.data_ptr(), b - a, copy)
            assert rc == 0, rc
            d_tape = torch.empty(tot[1] + 8, dtype=torch.int64, device=dev)
            d_str = torch.empty(tot[2] + 64, dtype=torch.uint8, device=dev)
            assert sp.emit(0, 0, 0, d_tape.data_ptr(), d_tape.numel(), d_str.data_ptr(), d_str.numel(), sp.bases_ptr) == 0
            rc, ex = sp.exchange_result()
            assert rc == 0 and ex[8] == 0 and ex[9] == it + 1, ex
        np.save(os.path.join(outdir, "tape%d.npy" % rank), d_tape[: tot[1]].cpu().numpy().view(np.uint64))
        with open(os.path.join(outdir, "str%d.bin" % rank), "wb") as f:
            f.write(d_str[: tot[2]].cpu().numpy().tobytes())
        np.save(os.path.join(outdir, "ex%d.npy" % rank), np.array(ex, dtype=np.uint64))
        dist.barrier()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("copy", [True, False])
def test_two_processes_exchange_through_cuda_ipc(tmp_path, oracle_native, copy):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), copy, str(tmp_path)), nprocs=world, join=True)
    if any((tmp_path / ("unsupported%d" % r)).exists() for r in range(world)):
        pytest.skip("this box does not let processes share device memory (CUDA IPC): bench.py then exchanges through NCCL")
    stream = _stream()
    rc, tape, strings, (off, ln) = oracle_native.parse(stream, ndjson=True, copy_strings=copy)
    assert rc == 0
    got = np.concatenate([np.load(tmp_path / ("tape%d.npy" % r)) for r in range(world)])
    gstr = b"".join(open(tmp_path / ("str%d.bin" % r), "rb").read() for r in range(world))
    assert len(got) == len(tape) and np.array_equal(got, tape)
    assert gstr == strings
    ex = np.load(tmp_path / "ex1.npy")
    assert int(ex[4]) == ln and int(ex[5]) == len(tape) and int(ex[6]) == len(strings)
