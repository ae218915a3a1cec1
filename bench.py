#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched Othello hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--games G]

Workload (config.workload = "config3_random_playout"): G = 2**20 lock-step random-playout games
PER GPU from the standard opening, every position and move written to HBM (SoA trajectory).  A
"step" is one launch of the playout kernel over one fresh batch of G games (new game ids every
step).  Games are independent, so N GPUs run N shards with no data-path collective ("weak").

    value   = positions/s = sum over ranks and steps of plies played / max-over-ranks device time
    e2e     = the same through the host-buffer C ABI (othello_playout_host): start positions come
              from pinned host memory, per-game results (plies, final position) go back to host
              memory, trajectories stay in HBM -- copies inside the timed region
    roofline= integer roofline of the playout kernel: 512 INT32 lane-ops per position-step
               against the integer peak measured live by csrc/peak.cu (ALU + FMA pipes
              co-issuing LOP3 + IMAD; the ALU-pipe-only peak is reported next to it), plus the HBM
              write-out rate (17 B per position) against MEASURED_PEAKS.json
    cpu_baseline = the reference's own board.py (oracle/_ref, py3 transcription) playing the same
              kind of games on all host cores for a bounded time (N=1, rank 0 only)

`--impl reference` times only that CPU path and prints the same JSON shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "othello_positions_per_sec_legalgen_step"
UNIT = "positions/s"
LANE_OPS_PER_POSITION = 512          # 256 64-bit bit-ops per legal-gen + flip
BYTES_PER_POSITION = 17              # black u64 + white u64 + move u8 written per ply
T_MAX = 120


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's board.py on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """plays seeded random games (bare loop: puttables -> choose -> put_s -> is_game_over) until the
    deadline; returns (plies, games).  Runs in a multiprocessing worker."""
    kind, seed, gid0, budget_s, max_games = args
    full_path = kind == "reference_full"
    t_end = time.perf_counter() + budget_s
    plies = games = 0
    if kind.startswith("reference"):
        from oracle import refshim, make_golden as mg
        rb = refshim.load().board
        while time.perf_counter() < t_end and games < max_games:
            B = rb.Board()
            key = mg.rng_key(seed, gid0 + games)
            t = 0
            while not B.is_game_over():
                moves = B.puttables(B.turn)
                if moves:
                    x, y = moves[mg.below(mg.rng_draw(key, t, 1), len(moves))]
                    B.put_s(B.handstr_from_coord(x, y))
                else:
                    B.put_s('ps')
                if full_path:
                    # what play_a_turn adds per ply (game_runner.py:154-163 + RedisRecorder.add,
                    # game_recorder.py:107-114): the printed board, the serialised record, its end flag
                    str(B)
                    {'book': B.serialize_board(), 'whosturn': B.serialize_turn(), 'turn': B.nturn,
                     'end': B.is_game_over()}
                t += 1
            plies += t
            games += 1
    else:
        from oracle import lib as orc
        while time.perf_counter() < t_end and games < max_games:
            r = orc.playout(seed, gid0 + games, 64, trajectory=False)
            plies += int(r['nplies'].sum())
            games += 64
    return plies, games


def cpu_baseline(budget_s=12.0, max_games_per_worker=1 << 30, seed=1, full_path_s=0.0):
    """reference CPU path on all host cores for ~budget_s seconds (+ full_path_s seconds of the
    play_a_turn-equivalent path, reported separately)."""
    import multiprocessing as mp
    from oracle import refshim
    kind = "reference" if refshim.available() else "port"
    if kind == "port":
        from oracle import lib as orc
        orc.build()
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, seed, w * (1 << 24), budget_s, max_games_per_worker) for w in range(cores)])
    dt = time.perf_counter() - t0
    plies = sum(r[0] for r in res)
    games = sum(r[1] for r in res)
    what = ("reference board.py (oracle/_ref py3 transcription)" if kind == "reference"
            else "oracle/othello_oracle.c port (reference not present on this box)")
    extra = {}
    if kind == "reference" and full_path_s > 0:
        t1 = time.perf_counter()
        with ctx.Pool(cores) as pool:
            res2 = pool.map(_cpu_worker, [("reference_full", seed, w * (1 << 24), full_path_s, max_games_per_worker)
                                          for w in range(cores)])
        dt2 = time.perf_counter() - t1
        extra = {"full_play_a_turn_path": {"value": sum(r[0] for r in res2) / dt2, "unit": UNIT,
                                           "games_per_s": sum(r[1] for r in res2) / dt2,
                                           "what": "adds str(board), recorder-style serialisation and its "
                                                   "is_game_over per ply (game_runner.py:154-163, "
                                                   "game_recorder.py:107-114), %.1f s" % dt2}}
    return {"value": plies / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "games_per_s": games / dt, **extra,
            "sample": "%d random-playout games (%d plies) from the standard opening, bare loop "
                      "puttables->choose->put_s->is_game_over, %s, multiprocessing.Pool(%d), %.1f s"
                      % (games, plies, what, cores, dt)}, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # each step = a bounded sample of the workload; whole run stays within a few minutes
    per_step = args.cpu_seconds if args.cpu_seconds else max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_baseline(budget_s=min(per_step, 3.0))
    total_plies = total_games = 0.0
    total_dt = 0.0
    last = None
    for _ in range(args.steps):
        last, dt = cpu_baseline(budget_s=per_step)
        total_plies += last["value"] * dt
        total_games += last["games_per_s"] * dt
        total_dt += dt
    value = total_plies / total_dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "games_per_s": total_games / total_dt,
        "config": {"workload": "config3_random_playout", "games_per_gpu": args.games, "t_max": T_MAX,
                   "note": "each step is a %.0f s time-bounded sample of the same workload on all host cores" % per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """samples SM clock and throttle reasons through NVML in a thread WHILE the timed region runs
    (the region lasts tens of ms, too short for `nvidia-smi -lms`); same fields as the recipe's
    nvidia-smi line: clocks.sm, clocks.max.sm, clocks_event_reasons.*"""

    def __init__(self, cuda_index):
        self.cuda_index = cuda_index
        self.samples = []
        self.reasons = 0
        self.max_mhz = None
        self.handle = None
        self.err = None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(self.cuda_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.cuda_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._sample()
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception as e:                       # pragma: no cover - depends on the box
            self.err = "%s: %s" % (type(e).__name__, e)
            self.handle = None

    def _sample(self):
        nv = self.nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))

    def _pump(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.001)

    def stop(self):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self._stop.set()
        self.thread.join(timeout=2)
        nv = self.nv
        names = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap),
                 ("hw_power_brake", nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown)]
        reasons = [n for n, bit in names if self.reasons & bit]
        sm = sorted(self.samples[1:] or self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "how": "NVML SM clock + throttle reasons polled every ~1 ms during the timed region"}


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200_arm(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from subproc_b200 import ops, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path is sm_100a kernels, there is no CPU fallback")
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        secs = args.cpu_seconds if args.cpu_seconds else 12.0
        cb, _ = cpu_baseline(budget_s=secs, full_path_s=secs / 2)          # before CUDA is initialised (fork-safe)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    G = args.games
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: value ---------------------------------------------------------
    po = ops.playout(G, seed=1, gid0=0, device=dev, t_max=T_MAX)          # allocates the 2 GB trajectory once
    nplies_steps = [torch.empty(G, dtype=torch.int32, device=dev) for _ in range(K)]
    gid_base = rank * (W + K) * G * 2

    def one_step(i, nplies_out=None):
        if nplies_out is not None:
            po.nplies = nplies_out
        ops.playout(G, seed=1, gid0=gid_base + i * G, device=dev, t_max=T_MAX, out=po)

    for i in range(W):
        one_step(i)
    int_peak_alu = ops.int32_peak(dev)                                     # ALU pipe alone (LOP3/SHF), lane-ops/s
    int_peak = ops.int32_peak(dev, dual=True)                              # ALU + FMA pipes (LOP3 + IMAD): the integer ceiling
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    evs[0].record()
    for i in range(K):
        one_step(W + i, nplies_steps[i])
        evs[i + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    total_ms = evs[0].elapsed_time(evs[K])
    positions = sum(int(t.sum(dtype=torch.int64).item()) for t in nplies_steps)
    games = K * G

    # ---- the batch API on explicit positions: othello_step (put_s + game-over check) on synthetic
    # random-opening positions = the positions of this launch after 10 random plies, their next move
    t_open = 10
    sb, sw = po.black[t_open].clone(), po.white[t_open].clone()
    mv = po.move[t_open].contiguous()
    wb, ww = torch.empty_like(sb), torch.empty_like(sw)
    turn = torch.empty(G, dtype=torch.uint8, device=dev)
    nturn = torch.empty(G, dtype=torch.int32, device=dev)
    fl, rt, fg = torch.empty_like(sb), torch.empty(G, dtype=torch.int32, device=dev), torch.empty(G, dtype=torch.uint8, device=dev)
    step_times = []
    for i in range(3 + 5):
        wb.copy_(sb); ww.copy_(sw); turn.fill_(ops.BLACK); nturn.fill_(t_open)      # restore (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.step(wb, ww, turn, nturn, mv, fl, rt, fg)
        e1.record()
        e1.synchronize()
        if i >= 3:
            step_times.append(e0.elapsed_time(e1))
    step_live = int((rt > 0).sum().item())
    step_info = {"positions_per_s": G / (min(step_times) * 1e-3), "kernel_ms": min(step_times), "positions": G,
                 "legal_moves_applied": step_live,
                 "what": "othello_step (put_s + pass / game-over flags) on the positions after %d random plies; "
                         "50 B of HBM traffic per position" % t_open}

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------
    L = _lib.lib()
    ctx = ctypes.c_void_p()
    _lib.check(L.othello_ctx_create(local, ctypes.byref(ctx)), "othello_ctx_create")
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_b0 = pin(G, torch.int64).fill_(ops.signed64(ops.START_BLACK))
    h_w0 = pin(G, torch.int64).fill_(ops.signed64(ops.START_WHITE))
    h_t0 = pin(G, torch.uint8).fill_(ops.BLACK)
    h_np, h_fb, h_fw = pin(G, torch.int32), pin(G, torch.int64), pin(G, torch.int64)
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    def e2e_step(i):
        _lib.check(L.othello_playout_host(ctx, 1, gid_base + (W + K + i) * G, G, P(h_b0), P(h_w0), P(h_t0),
                                          ops.POLICY_RANDOM, 0, 0, 0, None, -1, None, T_MAX, None, None, None,
                                          P(h_np), P(h_fb), P(h_fw)), "othello_playout_host")
        return int(h_np.sum(dtype=torch.int64).item())

    for i in range(max(1, min(W, 2))):
        e2e_step(K + i)
    barrier()
    t0 = time.perf_counter()
    e2e_positions = 0
    for i in range(K):
        e2e_positions += e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    L.othello_ctx_destroy(ctx)
    h2d = G * (8 + 8 + 1)
    d2h = G * (4 + 8 + 8)

    # ---- reduce over ranks -----------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([positions, games, e2e_positions], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms = float(t[0]), float(t[1])
        positions, games, e2e_positions = float(c[0]), float(c[1]), float(c[2])
    else:
        e2e_ms = e2e_s * 1e3

    if rank == 0:
        value = positions / (total_ms * 1e-3)
        hbm_peak, hbm_src = measured_peaks()
        # the dominant (only) kernel of a step is playout_kernel<false,false,true>: per launch
        kern_ms = sum(step_ms) / K
        pos_per_launch = positions / (K * n_gpus)
        achieved_ops = pos_per_launch * LANE_OPS_PER_POSITION / (kern_ms * 1e-3)
        achieved_gbs = pos_per_launch * BYTES_PER_POSITION / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "playout_traffic.json")
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "games_per_s": games / (total_ms * 1e-3),
            "config": {"workload": "config3_random_playout", "games_per_gpu": G, "t_max": T_MAX,
                       "policy": "uniform random, standard opening, seed 1",
                       "trajectory": "full SoA [t][game] black,white u64 + move u8 in HBM",
                       "l2": "kernel reads no input from HBM; ~%.2f GB written per step exceeds the 126 MB L2"
                             % (pos_per_launch * BYTES_PER_POSITION / 1e9),
                       "parallelism": "games sharded over %d GPU(s), no data-path collective" % n_gpus},
            "roofline": {"bound": "int32", "achieved": achieved_ops / 1e12, "peak": int_peak / 1e12,
                         "unit": "Tlane-op/s", "frac": achieved_ops / int_peak, "traffic": traffic,
                         "peak_source": "csrc/peak.cu, measured in this run: LOP3 + IMAD co-issued on the ALU and FMA "
                                        "pipes (the two pipes that execute 32-bit integer lane-ops)",
                         "peak_alu_pipe_only": int_peak_alu / 1e12, "frac_alu_pipe_only": achieved_ops / int_peak_alu,
                         "algorithmic": "%d INT32 lane-ops per position-step (SURVEY.md 8d)" % LANE_OPS_PER_POSITION,
                         "kernel": "playout_kernel<random,traj>", "kernel_ms": kern_ms,
                         "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src,
                                 "algorithmic": "%d B written per position" % BYTES_PER_POSITION}},
            "e2e": {"value": e2e_positions / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "games_per_s": G * K * n_gpus / (e2e_ms * 1e-3),
                    "api": "othello_playout_host (C ABI, pinned host buffers; trajectories stay in HBM)"},
            "gpu_launches": K,
            "step_kernel": step_info,
            "clocks": clocks,
        }
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=1 << 20, help="games per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=None,
                    help="seconds of host-core time per CPU sample (default: 12 for the cpu_baseline leg, "
                         "90 / (steps + warmup) clamped to [2, 20] per step of --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                      # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
