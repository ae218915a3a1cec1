#!/usr/bin/env python
"""tools/e2e_breakdown.py -- where the end-to-end (host-buffer) playout step spends its time.

    python tools/e2e_breakdown.py [--steps K] [--bind]                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/e2e_breakdown.py

Every phase is timed on all ranks at once (barrier before, wall clock after a full sync, max over
ranks), so contention between ranks for PCIe, host memory and host cores shows up:

    kernel      device-resident playout of G games (CUDA events)
    h2d / d2h   the copies of one step alone (17 B / 20 B per game), pinned buffers
    duplex      both directions at once on two streams
    host_sum    what round 1's bench loop did with the results on the CPU, three ways
    sync        othello_playout_host: one call = copy-in, kernels, copy-out, synchronise
    async       othello_playout_host_async with two batches in flight (+ variants: no upload of start
                positions, totals only, number of chunks)

Writes one JSON object (rank 0) to stdout and to gpurun_out/e2e_breakdown_n<N>.json.
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--games", type=int, default=1 << 20)
    ap.add_argument("--bind", action="store_true", help="give every rank its own share of the host cores")
    ap.add_argument("--threads", type=int, default=0, help="torch.set_num_threads (0 = leave the default)")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from bench import bind_rank_to_cores
    binding = bind_rank_to_cores(local, world) if args.bind else None

    import numpy as np
    import torch
    import torch.distributed as dist
    from subproc_b200 import ops, _lib
    if args.threads:
        torch.set_num_threads(args.threads)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    G, K = args.games, args.steps
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def wall(fn, reps=K, warm=2):
        """ms per repetition, max over ranks; fn(i) enqueues / runs repetition i"""
        for i in range(warm):
            fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(reps):
            fn(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return maxr(1e3 * dt / reps)

    out = {"n_gpus": world, "games_per_gpu": G, "steps": K, "binding": binding,
           "torch_threads": torch.get_num_threads(), "affinity": len(os.sched_getaffinity(0)), "tag": args.tag}

    # ---- kernel ------------------------------------------------------------------------------
    po = ops.playout(G, seed=1, gid0=0, device=dev)
    out["kernel_ms"] = wall(lambda i: ops.playout(G, seed=1, gid0=(rank * 1000 + i) * G, device=dev, out=po))
    positions_per_step = po.total_positions()

    # ---- raw copies --------------------------------------------------------------------------
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_b0 = pin(G, torch.int64).fill_(ops.signed64(ops.START_BLACK))
    h_w0 = pin(G, torch.int64).fill_(ops.signed64(ops.START_WHITE))
    h_t0 = pin(G, torch.uint8).fill_(ops.BLACK)
    bufs = [(pin(G, torch.int32), pin(G, torch.int64), pin(G, torch.int64), pin(4, torch.int64), pin(G, torch.int16))
            for _ in range(2)]
    d_b0, d_w0, d_t0 = h_b0.to(dev), h_w0.to(dev), h_t0.to(dev)
    d_np, d_fb, d_fw = po.nplies, po.final_black, po.final_white
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d(i):
        with torch.cuda.stream(s_in):
            d_b0.copy_(h_b0, non_blocking=True); d_w0.copy_(h_w0, non_blocking=True); d_t0.copy_(h_t0, non_blocking=True)

    def d2h(i):
        with torch.cuda.stream(s_out):
            bufs[0][0].copy_(d_np, non_blocking=True); bufs[0][1].copy_(d_fb, non_blocking=True)
            bufs[0][2].copy_(d_fw, non_blocking=True)

    out["h2d_ms"] = wall(h2d)
    out["d2h_ms"] = wall(d2h)
    out["duplex_ms"] = wall(lambda i: (h2d(i), d2h(i)))
    out["h2d_gbs"] = G * 17 / out["h2d_ms"] / 1e6
    out["d2h_gbs"] = G * 20 / out["d2h_ms"] / 1e6

    # ---- host-side reduction of the per-game results -------------------------------------------
    h_np = bufs[0][0]
    npv = h_np.numpy()
    out["host_sum_torch_ms"] = wall(lambda i: int(h_np.sum(dtype=torch.int64).item()))
    out["host_sum_numpy_ms"] = wall(lambda i: int(npv.sum(dtype=np.int64)))

    # ---- the C ABI -----------------------------------------------------------------------------
    ctx = ctypes.c_void_p()
    _lib.check(L.othello_ctx_create(local, ctypes.byref(ctx)), "othello_ctx_create")
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    gid = [rank * 100000 * G]

    def sync_call(i, reduce=None):
        b = bufs[0]
        _lib.check(L.othello_playout_host(ctx, 1, gid[0], G, P(h_b0), P(h_w0), P(h_t0), 0, 0, 0, 0, None, -1, None, 120,
                                          None, None, None, P(b[0]), P(b[1]), P(b[2])), "playout_host")
        gid[0] += G
        if reduce == "torch":
            return int(b[0].sum(dtype=torch.int64).item())
        if reduce == "numpy":
            return int(b[0].numpy().sum(dtype=np.int64))

    out["sync_ms"] = wall(sync_call)
    out["sync_plus_torch_sum_ms"] = wall(lambda i: sync_call(i, "torch"))       # round 1's bench loop
    out["sync_plus_numpy_sum_ms"] = wall(lambda i: sync_call(i, "numpy"))

    def async_loop(reps, upload=True, per_game=True, totals=True, summary=False, turn=True):
        """two batches in flight: issue i+1, wait i.  returns ms per batch (this rank)"""
        tickets = [None, None]
        total = 0
        t0 = time.perf_counter()
        for i in range(reps + 1):
            if i < reps:
                b = bufs[i % 2]
                tk = ctypes.c_int64()
                _lib.check(L.othello_playout_host_async(
                    ctx, 1, gid[0], G, P(h_b0) if upload else None, P(h_w0) if upload else None,
                    P(h_t0) if upload and turn else None, 0, 0, 0, 0, None, -1, None, 120, None, None, None,
                    P(b[0]) if per_game else None, P(b[1]) if per_game else None, P(b[2]) if per_game else None,
                    P(b[4]) if summary else None, P(b[3]) if totals else None, ctypes.byref(tk)), "playout_host_async")
                gid[0] += G
                tickets[i % 2] = tk.value
            if i > 0:
                _lib.check(L.othello_ctx_wait(ctx, tickets[(i - 1) % 2]), "ctx_wait")
                total += int(bufs[(i - 1) % 2][3][0]) if totals else 0
        return 1e3 * (time.perf_counter() - t0) / reps, total

    def timed_async(**kw):
        async_loop(3, **kw)
        barrier()
        ms, total = async_loop(K, **kw)
        return maxr(ms), total

    out["async_ms"], tot = timed_async()
    out["async_positions_per_step"] = tot / K
    out["async_no_upload_ms"], _ = timed_async(upload=False)
    out["async_totals_only_ms"], _ = timed_async(upload=False, per_game=False)
    out["async_upload_summary_ms"], _ = timed_async(upload=True, per_game=False, summary=True, turn=False)   # 16 B in, 2 B out per game
    out["async_summary_only_ms"], _ = timed_async(upload=False, per_game=False, summary=True)
    for ch in (1, 2, 4, 16):
        _lib.check(L.othello_ctx_set_option(ctx, 1, ch), "set_option")
        out["async_chunks%d_ms" % ch], _ = timed_async()
    L.othello_ctx_destroy(ctx)
    out["positions_per_step"] = positions_per_step
    out["e2e_over_kernel"] = out["kernel_ms"] / out["async_ms"]

    if rank == 0:
        s = json.dumps(out, indent=1)
        print(s)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "e2e_breakdown_n%d%s.json" % (world, args.tag)), "w") as f:
            f.write(s + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
