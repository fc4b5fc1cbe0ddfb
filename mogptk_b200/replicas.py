"""Outer-level multi-GPU plumbing: independent replicas, one process per GPU.

The exact-GP step is a dense factorisation with a sequential panel dependency and is not
sharded ("replicas only").  What scales over the 8xB200 box is the loop one level
up that the reference runs serially -- random restarts, channel-subset models, hyper-parameter
sweeps (e.g. examples/example_currency_exchange.ipynb trains 3 restarts x 4 models).  Each rank
trains its own replica; the only communication is one all-gather of the final losses (8 bytes per
rank) and, optionally, a broadcast of the winner's parameter vector.  NCCL over NVLink on GPUs,
gloo in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None, device=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    rank, world, local = env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = device if device is not None else torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def _dev(device):
    return device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                              if torch.cuda.is_available() and dist.is_initialized()
                                              and dist.get_backend() == "nccl" else torch.device("cpu"))


def all_gather_scalar(x, device=None):
    """Every rank's value of x (one fp64 per rank -- the single collective of the path)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(x)]
    d = _dev(device)
    mine = torch.tensor([float(x)], dtype=torch.float64, device=d)
    out = [torch.zeros(1, dtype=torch.float64, device=d) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [float(o.item()) for o in out]


def max_over_ranks(x, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_dev(device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def best_replica(losses):
    """Index of the replica with the lowest finite loss (ties -> lowest rank)."""
    best, arg = float("inf"), -1
    for i, l in enumerate(losses):
        if l == l and l < best:
            best, arg = l, i
    return arg


def broadcast_winner(packed, losses, device=None):
    """Give every rank the parameter vector of the best replica."""
    src = best_replica(losses)
    if not dist.is_initialized() or dist.get_world_size() == 1 or src < 0:
        return packed, src
    t = packed.detach().clone().to(_dev(device))
    dist.broadcast(t, src=src)
    return t.to(packed.device), src


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def finish():
    if dist.is_initialized():
        dist.destroy_process_group()


# ------------------------------------------------------------------ several replicas on ONE GPU
# At BASELINE configs[1] (N = 2048) one evaluation is a latency chain: the Cholesky panel kernel occupies 46 of the 148
# SMs and the GPU draws ~270 W of 1000 W.  Independent replicas (restarts, sweeps) therefore also stack on one GPU: each
# gets its own workspace handle and CUDA stream, and their kernels interleave on the idle SMs.
def train_restarts(models, iters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, sync_every=64):
    """Train independent mogptk_b200.gpr.Exact models concurrently on the current GPU with the device-resident Adam loop
    (one host thread and one CUDA stream per model).  Returns the list of per-model loss histories."""
    import threading

    from .train import fit_adam
    out = [None] * len(models)
    errs = []

    def work(i, m):
        try:
            stream = torch.cuda.Stream(device=m.X.device)
            with torch.cuda.stream(stream):
                out[i] = fit_adam(m, iters, lr=lr, betas=betas, eps=eps, sync_every=sync_every)[0]
                stream.synchronize()
        except BaseException as e:          # surfaced to the caller below
            errs.append(e)

    torch.cuda.synchronize()
    threads = [threading.Thread(target=work, args=(i, m)) for i, m in enumerate(models)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return out
