"""CPU, world_size = 2 over gloo: the replica plumbing (the only multi-GPU logic of the path)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from mogptk_b200 import replicas
    r, w, _ = replicas.init(backend="gloo")
    loss = 10.0 - 3.0 * r                      # rank 1 "wins"
    losses = replicas.all_gather_scalar(loss)
    tmax = replicas.max_over_ranks(1.0 + r)
    packed = torch.full((5,), float(r), dtype=torch.float64)
    win, src = replicas.broadcast_winner(packed, losses)
    replicas.barrier()
    q.put((r, w, losses, tmax, win.tolist(), src))
    replicas.finish()


def test_two_rank_replicas_exchange_only_their_losses():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r, (rank, world, losses, tmax, win, src) in enumerate(res):
        assert rank == r and world == 2
        assert losses == [10.0, 7.0]
        assert tmax == 2.0
        assert src == 1 and win == [1.0] * 5


def test_single_process_is_a_no_op():
    from mogptk_b200 import replicas
    assert replicas.all_gather_scalar(3.5) == [3.5]
    assert replicas.max_over_ranks(2.0) == 2.0
    assert replicas.best_replica([float("nan"), 4.0, 2.0, 2.0]) == 2
