"""N > 1 host logic on CPU: two gloo ranks. No GPU: only the host side of the sharded proof -- the gather of the
per-rank message buffers and their field sum (hg_shard_merge, a host function of the C-ABI library) -- and the max-over-ranks
timing reduction bench.py uses are exercised."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GL_P = (1 << 64) - (1 << 32) + 1
FR_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, field, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        import hyper_greco_b200  # noqa: F401
        from hyper_greco_b200 import api

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        n_el = 1000
        words = 2 if field == 0 else 4
        rng = np.random.default_rng(100 + rank)
        if field == 0:
            part = rng.integers(0, GL_P, size=n_el * words, dtype=np.uint64)
            part[::7] = GL_P - 1  # force the modular wrap
        else:
            vals = [int.from_bytes(rng.bytes(32), "little") % FR_R for _ in range(n_el)]
            vals[::5] = [FR_R - 1] * len(vals[::5])
            part = np.array([(v >> (64 * k)) & (2**64 - 1) for v in vals for k in range(4)], dtype=np.uint64)
        # slots owned by the other rank stay zero, as LassoNode.prove_shard leaves them
        owned = np.arange(n_el) % world == rank
        shared = np.arange(n_el) % 11 == 0
        mask = np.repeat(owned | shared, words)
        part = np.where(mask, part, np.uint64(0))
        merged = api.gather_and_merge(field, part, None)
        # bench.py's timing reduction: max over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, part, merged, float(t[0])))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "error", traceback.format_exc(), repr(e)))


@pytest.mark.parametrize("field", [0, 1])
def test_two_rank_gloo_gather_and_field_merge(field):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, field, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        item = q.get(timeout=180)
        assert item[1] is not None and not (isinstance(item[1], str) and item[1] == "error"), item[2]
        res[item[0]] = item
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[1][2] is None                      # only rank 0 holds the merged buffer
    assert res[0][3] == 11.0 and res[1][3] == 11.0  # max over ranks
    a, b, merged = res[0][1], res[1][1], res[0][2]
    if field == 0:
        want = np.array([(int(x) + int(y)) % GL_P for x, y in zip(a, b)], dtype=np.uint64)
    else:
        def ints(w):
            return [sum(int(w[4 * i + k]) << (64 * k) for k in range(4)) for i in range(w.size // 4)]
        s = [(x + y) % FR_R for x, y in zip(ints(a), ints(b))]
        want = np.array([(v >> (64 * k)) & (2**64 - 1) for v in s for k in range(4)], dtype=np.uint64)
    assert (merged == want).all()


def test_shard_merge_rejects_mismatched_buffers():
    sys.path.insert(0, ROOT)
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import api

    with pytest.raises(api.HgError):
        api.shard_merge(0, np.zeros(4, np.uint64), np.zeros(6, np.uint64))
    with pytest.raises(api.HgError):
        api.shard_merge(0, np.zeros(3, np.uint64), np.zeros(3, np.uint64))  # not a whole number of extension elements
