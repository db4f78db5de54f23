"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: the data-parallel gradient bucket of PCNet training
(train_network.shard_indices / allreduce_grads: global-batch semantics) and the collective-free attack-job sharder."""
import os
import random
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _toy_grad(w, xs):
    """Gradient of mean_i 0.5*(w.x_i)^2 over the samples xs (stands for one PCNet backward on a shard)."""
    return (xs @ w).unsqueeze(1).mul(xs).mean(0)


class _Bucket:                      # the two FlatAdam fields allreduce_grads touches
    def __init__(self, n):
        self.grad = torch.zeros(n)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spaa_b200 import train_network as tn
        from spaa_b200.projector_based_attack import gather_job_results, shard_jobs
        assert tn._world() == (rank, world)
        # ---- training: every rank draws the SAME seeded global batch and takes a strided share (train_network.py:295) ----
        torch.manual_seed(0)
        data, w = torch.randn(500, 16), torch.randn(16)
        random.seed(123)
        errs = []
        for step in range(3):
            idx = random.sample(range(500), 24)
            mine = tn.shard_indices(idx, rank, world)
            assert len(mine) == 24 // world
            opt = _Bucket(16)
            opt.grad.copy_(_toy_grad(w, data[mine]))
            scale = tn.allreduce_grads(opt, world)
            got = opt.grad * scale                                   # what spaa_adam_step consumes (grad_scale)
            want = _toy_grad(w, data[idx])                           # single-process global-batch gradient
            errs.append((got - want).abs().max().item())
            w = w - 0.1 * got
        # shards of one draw are disjoint and cover it
        cover = [None] * world
        dist.all_gather_object(cover, mine)
        assert sorted(sum(cover, [])) == sorted(idx)
        # ---- attack sweep: 3 classifiers x 5 setups, independent jobs, no collective on the data path ----
        jobs = [(c, s) for c in ("resnet18", "vgg16", "inception_v3") for s in range(5)]
        local = [(i, f"{j[0]}:{j[1]}:done-by-{rank}") for i, j in shard_jobs(jobs)]
        allr = gather_job_results(local, len(jobs))
        assert [r.rsplit(":", 1)[0] for r in allr] == [f"{c}:{s}" for c, s in jobs]
        assert sum(r.endswith(f"-{rank}") for r in allr) == len(local)
        ret[rank] = max(errs)
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_bucket_and_job_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
    assert len(ret) == world and max(ret.values()) < 1e-6, dict(ret)


def test_sharding_single_process_defaults():
    from spaa_b200 import train_network as tn
    from spaa_b200.projector_based_attack import gather_job_results, shard_jobs
    assert tn._world() == (0, 1)
    assert tn.shard_indices(list(range(10)), 1, 4) == [1, 5, 9]
    jobs = list("abcde")
    assert shard_jobs(jobs) == list(enumerate(jobs))
    assert shard_jobs(jobs, 1, 2) == [(1, "b"), (3, "d")]
    assert gather_job_results([(i, j.upper()) for i, j in shard_jobs(jobs)], 5) == list("ABCDE")


def test_shard_jobs_cost_balanced_partition_is_deterministic_and_complete():
    """The cost-aware job partition (longest-processing-time-first) of an attack sweep: every rank computes it locally, the shares are disjoint,
    cover the sweep and are better balanced than round-robin for the 3-classifier cost mix."""
    from spaa_b200.projector_based_attack import shard_jobs
    jobs = [(c, s) for s in range(10) for c in ("resnet18", "vgg16", "inception_v3")]
    cost = {"resnet18": 1.0, "inception_v3": 2.2, "vgg16": 2.8}
    costs = [cost[c] for c, _ in jobs]
    for world in (1, 2, 4, 8):
        parts = [shard_jobs(jobs, r, world, costs) for r in range(world)]
        idx = sorted(i for p in parts for i, _ in p)
        assert idx == list(range(len(jobs)))
        assert parts == [shard_jobs(jobs, r, world, costs) for r in range(world)]
        load = [sum(costs[i] for i, _ in p) for p in parts]
        rr = [sum(costs[i] for i in range(r, len(jobs), world)) for r in range(world)]
        assert max(load) <= max(rr) + 1e-9 and max(load) - min(load) <= 2.8 + 1e-9, (world, load, rr)
    assert shard_jobs(jobs, 1, 4) == [(i, jobs[i]) for i in range(1, 30, 4)]            # default stays round-robin
