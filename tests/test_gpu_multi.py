"""Multi-GPU path on a box with >= 2 B200s: one process per GPU over NCCL.  Games shard
by global game id with no data-path collective; the learner's only exchange is one all-reduce of the
[4][112] statistics.  Skipped on single-GPU boxes (the CPU suite covers the same logic over gloo)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WORLD = 2


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from subproc_b200 import ops, learner, parameter
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    total = 8192
    lo, hi = learner.shard_of_games(total, rank, world)
    w = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(dev)
    po = ops.playout(hi - lo, seed=17, gid0=lo, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w)
    acc = ops.learn_accumulate(po)
    L = learner.ProgressPositionMovesLearn()
    rows = L.learn_from_acc(acc)                                 # all-reduce of the int64 accumulators inside
    nodes = ops.perft_distributed(10, device=dev)               # depth-first stage split over the ranks, one u64 all-reduce
    q.put((rank, po.nplies.cpu().numpy(), ops.bits_numpy(po.final_black), acc.cpu().numpy(), L.read_parameters(), nodes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason="needs >= 2 GPUs")
def test_two_ranks_equal_one_gpu_on_the_union_of_games():
    import torch.multiprocessing as mp
    from subproc_b200 import ops, learner, parameter
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, WORLD, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in range(WORLD)), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    w = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(dev)
    po = ops.playout(8192, seed=17, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w)
    assert np.array_equal(np.concatenate([g[1] for g in got]), po.nplies.cpu().numpy())          # same games, sharded
    assert np.array_equal(np.concatenate([g[2] for g in got]), ops.bits_numpy(po.final_black))
    whole = ops.learn_accumulate(po)
    for g in got:                                                  # every rank holds the all-reduced accumulators:
        assert np.array_equal(g[3], whole.cpu().numpy())           # exact integer sums, the same bits as one GPU
    assert got[0][4] == got[1][4]                                  # identical parameters on all ranks
    assert got[0][5] == got[1][5] == 24571284                      # perft(10), SURVEY.md section 4
    L = learner.ProgressPositionMovesLearn()
    L.learn_from_acc(whole)
    assert L.read_parameters() == got[0][4]                        # ... and identical to the single-GPU fit


def _table_worker(rank, world, port, q):
    import torch.distributed as dist
    from subproc_b200 import ops, value_table
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    vt = value_table.ValueTable(device=dev)
    per = 3000
    for batch in range(2):                                        # the sharded table carries over between batches
        lo = (batch * world + rank) * per                         # contiguous ascending blocks of game ids per rank
        po = ops.playout(per, seed=23, gid0=lo, device=dev)
        keys, targets = vt.records_from_playout(po)
        vt.update_sharded(keys, targets)
    q.put((rank, vt.keys.cpu().numpy(), vt.values.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason="needs >= 2 GPUs")
def test_key_sharded_value_table_equals_the_single_gpu_table():
    """the exact value table sharded by key over the ranks, one all_to_all of records per batch;
    the union of the ranks' shares must be the single-GPU table, bit for bit"""
    import torch.multiprocessing as mp
    from subproc_b200 import ops, value_table
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_table_worker, args=(r, WORLD, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(WORLD)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    vt = value_table.ValueTable(device=dev)
    for batch in range(2):
        vt.update_from_playout(ops.playout(3000 * WORLD, seed=23, gid0=batch * WORLD * 3000, device=dev))
    want = dict(zip(vt.keys.cpu().numpy().tolist(), vt.values.cpu().numpy().tolist()))
    union = {}
    for _, k, v in got:
        assert not (set(k.tolist()) & set(union))                 # shares are disjoint
        union.update(zip(k.tolist(), v.tolist()))
    assert union == want
