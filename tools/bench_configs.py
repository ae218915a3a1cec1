#!/usr/bin/env python
"""tools/bench_configs.py -- BASELINE configs 4 and 5 (greedy self-play, parallel learner).

    python tools/bench_configs.py --workload selfplay [--games 524288] [--steps 5] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_configs.py --workload learner --games 65536

selfplay (config 4): every rank plays `games` greedy self-play games per step (first 10 plies uniformly
random, then arg-max of the linear evaluation on the default_value() rows), trajectories to HBM.
learner (config 5): per step = one learning iteration: greedy self-play with the current weights,
othello_learn_accumulate, ONE all-reduce of 4x80 int64 accumulators over NCCL, four 10x10 solves, +-127 scaling,
int() truncation.  Timing: CUDA events, max over ranks; rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from subproc_b200 import ops, learner, parameter

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=["selfplay", "learner"], default="selfplay")
    ap.add_argument("--games", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--random-plies", type=int, default=10)
    ap.add_argument("--device-solve", action="store_true", help="learner: refit on the device, no host round trip")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    G = args.games or (1 << 19 if args.workload == "selfplay" else 1 << 16)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    P = parameter.ProgressPositionMovesParameter()
    w = torch.from_numpy(P.weights_table()).to(dev)
    po = ops.playout(G, seed=2, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=args.random_plies, weights=w)
    positions = 0
    allreduce_ms = []
    L = learner.ProgressPositionMovesLearn()
    L.configure({})

    def step(i):
        nonlocal positions, w
        gid0 = ((i * world) + rank) * G
        if args.workload == "selfplay":
            ops.playout(G, seed=2, gid0=gid0, device=dev, policy=ops.POLICY_GREEDY, random_plies=args.random_plies,
                        weights=w, out=po)
            return None
        w = torch.from_numpy(L.weights_table()).to(dev)
        ops.playout(G, seed=2, gid0=gid0, device=dev, policy=ops.POLICY_GREEDY, random_plies=args.random_plies,
                    weights=w, out=po)
        acc = L.accumulate(po)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        learner.allreduce_stats(acc)                         # 320 int64: exact sums
        e1.record()
        stats = ops.learn_stats(acc)
        L.last_stats = stats
        L.last_fits = learner.fit_from_stats(stats)          # D2H of 3.6 KB + four 10x10 solves on the host
        rows = [learner.scale_param(f['coef']) if f['n'] else tuple(L.read_parameters()[1 + 9 * s:10 + 9 * s])
                for s, f in enumerate(L.last_fits)]
        L.params = learner.stored_parameters(rows)
        return (e0, e1)

    nplies_sum = torch.zeros((), dtype=torch.int64, device=dev)
    if args.workload == "learner" and args.device_solve:
        # the whole loop is enqueued without waiting for the host: K iterations in one call
        L.self_play_iterations_on_device(W, G, seed=2, first_iteration=0, random_plies=args.random_plies, device=dev,
                                         rank=rank, world=world)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        L.self_play_iterations_on_device(K, G, seed=2, first_iteration=W, random_plies=args.random_plies, device=dev,
                                         rank=rank, world=world)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"workload": "config5_parallel_learner_device_solve", "n_gpus": world,
                              "games_per_gpu_per_step": G, "steps": K, "warmup": W, "ms_per_step": float(t[0]) / K,
                              "games_per_s": K * G * world / (float(t[0]) * 1e-3),
                              "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / K,
                              "parameters": list(L.read_parameters())}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    for i in range(W):
        step(i)
        nplies_sum += po.nplies.sum(dtype=torch.int64)     # also warms torch's lazily loaded reduce kernel
    barrier()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nplies_sum.zero_()
    ev0.record()
    pend = []
    for i in range(K):
        r = step(W + i)
        nplies_sum += po.nplies.sum(dtype=torch.int64)
        if r:
            pend.append(r)
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = ev0.elapsed_time(ev1)
    allreduce_ms = [a.elapsed_time(b) for a, b in pend]
    t = torch.tensor([ms, wall_ms], dtype=torch.float64, device=dev)
    c = torch.tensor([float(nplies_sum.item()), float(K * G)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        params = torch.tensor(L.read_parameters(), dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(params) for _ in range(world)]
        dist.all_gather(gathered, params)
        same = all(torch.equal(g, params) for g in gathered)
    else:
        same = True
    if rank == 0:
        ms, wall_ms = float(t[0]), float(t[1])
        line = {"workload": "config4_greedy_selfplay" if args.workload == "selfplay" else "config5_parallel_learner",
                "n_gpus": world, "games_per_gpu_per_step": G, "steps": K, "warmup": W,
                "random_plies": args.random_plies,
                "positions_per_s": float(c[0]) / (ms * 1e-3), "games_per_s": float(c[1]) / (ms * 1e-3),
                "ms_per_step": ms / K, "wall_ms_per_step": wall_ms / K, "timing": "CUDA events, max over ranks"}
        if args.workload == "learner":
            line["allreduce_ms"] = sum(allreduce_ms) / max(1, len(allreduce_ms))
            line["allreduce_bytes"] = 4 * 112 * 8
            line["params_identical_on_all_ranks"] = bool(same)
            line["parameters"] = list(L.read_parameters())
            line["fits"] = [{"n": f["n"], "rmse": f["rmse"], "r2": f["r2"]} for f in L.last_fits]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
