#!/usr/bin/env python
"""bench.py -- ray-samples/s through the sampling + compositing hot path.

Metric (BASELINE.json): samples/s through `OccGridEstimator.sampling` (grid traversal)
+ `rendering` (render_weight_from_density + 3x accumulate_along_rays) forward AND
backward, 65 536 rays on a 128^3 occupancy grid, ~128 samples/ray, per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this implementation
    python bench.py --impl reference ...                        # CPU arm (see below)
    torchrun --nproc-per-node N bench.py --gpus N ...           # weak scaling, rays sharded

One step = one pass of the hot path over one batch of synthetic rays:
sampling -> rendering -> MSE loss on colours -> backward (d/dsigmas, d/drgbs).  sigmas /
rgbs stand in for the user's radiance field (seeded leaf tensors).

`value`   samples/s with the ray batch already resident in HBM.
`e2e`     same step through the public API with HOST inputs: every step copies its ray batch
          from pinned host memory (double-buffered: the copy for step k+1 is started while step k
          waits for its march) and one scalar loss goes back to the host (read one step late, as
          a logger would), all inside the timed region.
N > 1     ray-sharded weak scaling; the one exchange of the path, the sum of the scalar loss,
          goes through nerfacc_b200.parallel (NVLink peer mailbox, NCCL fallback), started
          before backward() and consumed a step later.
clocks    `nvidia-smi -lms 200` runs from program start; the value arm puts CLOCK_LOAD_STEPS
          extra untimed steps (on every rank) in front of its W warm-up steps so that samples
          fall under load without leaving an idle gap in front of the timed region.
`--impl reference`  the reference has no CPU implementation of the packed path
          (nerfacc/pack.py:47-48); the arm therefore times the CPU oracle port
          (oracle/oracle.c, OpenMP over rays) of exactly this path on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ray-samples/sec (traverse+composite fwd+bwd)"
RAYS_PER_GPU = 65536
GRID_RES = 128

def workload_name() -> str:
    """The one workload both arms run (the driver compares the strings): config 2 of BASELINE.json."""
    return (f"{GRID_RES}^3 occ-grid (ball, 6.5% occupied), {RAYS_PER_GPU} rays/GPU, ~130 samples/ray, "
            "traverse + composite fwd+bwd")


# algorithmic bytes (SURVEY.md 8d / DESIGN.md "Roofline")
B_TRAVERSE, B_FWD, B_BWD, B_RAY = 16, 44, 48, 88
LOSS_LAG = int(os.environ.get("NFA_BENCH_LOSS_LAG", "2"))  # steps between starting the loss all-reduce and consuming its result (N > 1)
CLOCK_LOAD_STEPS = int(os.environ.get("NFA_BENCH_CLOCK_LOAD_STEPS", "1500"))  # ~0.5 s of untimed steps in front of the value arm so nvidia-smi samples fall under load
COLLECTIVE = os.environ.get("NFA_BENCH_NO_COLLECTIVE", "0") == "0"  # debugging aid: time N ranks without the all-reduce
LOSS_TRANSPORT = os.environ.get("NFA_BENCH_LOSS_TRANSPORT", "peer")  # "peer": NVLink mailbox (csrc/peer.cu); "nccl": dist.all_reduce
NOWAIT = os.environ.get("NFA_BENCH_LOSS_NOWAIT", "0") != "0"  # debugging aid: never consume the reduced loss inside the loop
LOSS_DEFER = os.environ.get("NFA_BENCH_LOSS_DEFER", "1") != "0"  # park its host-side enqueue in the next march wait


def load_traffic(kernel, n_samples):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r2_traffic.json), or None when
    the capture was taken on a different sample count."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if abs(t["n_samples"] - n_samples) > 0.01 * n_samples or kernel not in t:
            return None
        return t[kernel]["dram_read"] + t[kernel]["dram_write"]
    except (OSError, ValueError, KeyError):
        return None


def pin_to_gpu_numa_node(index: int) -> None:
    """Run this process on the CPUs next to its GPU (one process per GPU; torchrun does not pin).  The step is
    host-bound: launches and pinned-memory copies from the far socket cost tens of microseconds per step."""
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev_id = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except (OSError, ValueError, AttributeError):
        pass


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled under the benchmark's load.

    One background `nvidia-smi -lms 200` process (the profiling recipe's clocks line) started when the program
    starts and stopped right after the timed region.  The step is partly host-bound, so nothing here runs
    Python, sleeps or spawns processes near the timed region: an idle gap in front of it would let the GPU
    drop its clocks, and a 30-step region lasts ~10 ms, shorter than any sampling period -- so the value arm
    runs `CLOCK_LOAD_STEPS` extra untimed steps (on every rank) in front of the region and the samples kept are
    the ones taken during that load and the region itself.
    """

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/nfa_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu=timestamp,{self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_load(self):
        """Samples from now on are taken under load."""
        self.t_load = time.time()

    def stop(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.out.close()
            try:
                import datetime
                for l in open(self.path):
                    c = [x.strip() for x in l.split(",")]
                    if len(c) < 7:
                        continue
                    try:  # "2026/09/22 21:54:03.123"
                        ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    except ValueError:
                        continue
                    if ts >= getattr(self, "t_load", 0.0) + 0.05:
                        rows.append(c[1:])
                os.remove(self.path)
            except Exception:
                pass
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def physical_cores(orc) -> int:
    """Threads for the OpenMP oracle: one per physical core this process may use (SMT siblings only add
    contention to these memory-bound loops: 128 threads ran 1.3x slower than 64 on the 2 x 32-core hosts)."""
    logical = orc.host_cores()
    try:
        import psutil
        phys, total = psutil.cpu_count(logical=False), psutil.cpu_count(logical=True)
        if phys and total and total > phys:
            return max(1, logical * phys // total)
    except Exception:
        pass
    return logical


def cpu_oracle_step(orc, ro, rd, bins, aabbs, step_size, sig_seed=43):
    """One pass of the hot path on the host cores with the oracle port. Returns (n_samples, seconds)."""
    t0 = time.perf_counter()
    ri, ts, te, pi = orc.occgrid_sampling(ro, rd, bins, aabbs, render_step_size=step_size)
    n = len(ri)
    rng = np.random.default_rng(sig_seed)
    t_gen = time.perf_counter()
    sig = (5 * rng.random(n)).astype(np.float32)
    rgb = rng.random((n, 3)).astype(np.float32)
    t_gen = time.perf_counter() - t_gen  # synthetic field values are not part of the path
    o = orc.composite(ts, te, sig, rgb, packed_info=pi)
    gC = (2.0 / o["colors"].size) * (o["colors"] - 0.5)
    orc.composite_backward(ts, te, sig, rgb, pi, gC=gC.astype(np.float32))
    return n, time.perf_counter() - t0 - t_gen


def bench_config(n_samples: int, step_size: float, world: int) -> dict:
    """`config` of the JSON line -- identical in both arms (the driver compares them)."""
    return {"workload": workload_name(), "samples_per_ray": round(n_samples / RAYS_PER_GPU, 1),
            "n_samples_per_gpu": int(n_samples), "render_step_size": step_size, "parallelism": f"ray-shard dp{world}"}


def _import_reference():
    """The UNMODIFIED reference package installed under baseline/_ref (pip --target), or None.
    It shadows nothing of ours: bench.py imports `nerfacc_b200`, never the `nerfacc` alias."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "nerfacc")):
        return None, "baseline/_ref is not installed"
    if "nerfacc" in sys.modules and ref_dir not in (getattr(sys.modules["nerfacc"], "__file__", "") or ""):
        return None, "another `nerfacc` module is already imported"
    sys.path.insert(0, ref_dir)
    try:
        import nerfacc as ref
        assert ref_dir in ref.__file__
        return ref, None
    except Exception as ex:  # missing dependency, ABI mismatch of the prebuilt csrc.so, ...
        return None, f"{type(ex).__name__}: {ex}"
    finally:
        sys.path.remove(ref_dir)


def time_reference_cuda(dev, ro_d, rd_d, step_size, R, N_ours, iters=10, warmup=3):
    """the reference's own CUDA build, same scene, same step (sampling + rendering + MSE +
    backward) through ITS public API, in this process right after our timed region; CUDA events."""
    ref, why = _import_reference()
    if ref is None:
        return {"unavailable": why}
    try:
        from nerfacc_b200 import scenes
        est = ref.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=GRID_RES).to(dev)
        est.binaries = torch.from_numpy(scenes.ball_grid(GRID_RES)).to(dev)
        ri, ts, te = est.sampling(ro_d, rd_d, render_step_size=step_size)
        N = ri.numel()
        g = torch.Generator(device="cpu").manual_seed(43)
        sig = (5 * torch.rand(N, generator=g)).to(dev).requires_grad_(True)
        rgb = torch.rand(N, 3, generator=g).to(dev).requires_grad_(True)
        tgt = torch.rand(R, 3, generator=g).to(dev)

        def sampling():
            return est.sampling(ro_d, rd_d, render_step_size=step_size)

        def render(ri_, ts_, te_):
            col, _, _, _ = ref.rendering(ts_, te_, ri_, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
            sig.grad = None
            rgb.grad = None
            torch.nn.functional.mse_loss(col, tgt).backward()

        def timed(fn):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        ms_step = timed(lambda: render(*sampling()))
        ms_samp = timed(sampling)
        ms_rend = timed(lambda: render(ri, ts, te))
        return {"impl": "reference CUDA build (baseline/_ref, unmodified, its own public API)", "n_samples": N,
                "same_samples_as_ours": bool(N == N_ours), "ms_per_step": ms_step, "value": N / (ms_step * 1e-3),
                "unit": "samples/s", "stages_us": {"sampling": round(ms_samp * 1e3, 1),
                                                   "rendering_fwd_bwd": round(ms_rend * 1e3, 1)},
                "steps": iters, "warmup": warmup}
    except Exception as ex:
        return {"unavailable": f"{type(ex).__name__}: {ex}"}


def time_reference_cpu_torch(n_rays=4096, n_samples=128, iters=20, warmup=3):
    """BASELINE.json config 1 / BASELINE.md B2: the reference's pure-PyTorch CPU path -- batched [4096,128]
    `rendering` forward + backward on host tensors, uniform sigma = 5 -- on this box's host cores."""
    ref, why = _import_reference()
    if ref is None:
        return {"unavailable": why}
    try:
        import nerfacc.volrend as rv
        rv.is_cub_available = lambda: True  # reach the batched branch without touching the CUDA module (B2)
        from nerfacc_b200 import scenes
        threads_before = torch.get_num_threads()
        cores = len(os.sched_getaffinity(0))
        ro, rd = scenes.ball_rays(n_rays)
        # uniform bins over each ray's chord through the ball of radius 0.5
        o, d = torch.from_numpy(ro).double(), torch.from_numpy(rd).double()
        bq = (o * d).sum(-1)
        disc = (bq * bq - ((o * o).sum(-1) - 0.25)).clamp_min(0).sqrt()
        t0, t1 = (-bq - disc), (-bq + disc)
        edges = t0[:, None] + (t1 - t0)[:, None] * torch.linspace(0, 1, n_samples + 1, dtype=torch.float64)[None]
        t_starts, t_ends = edges[:, :-1].float().contiguous(), edges[:, 1:].float().contiguous()
        g = torch.Generator().manual_seed(44)
        sig = torch.full((n_rays, n_samples), 5.0, requires_grad=True)
        rgb = torch.rand(n_rays, n_samples, 3, generator=g).requires_grad_(True)

        def step():
            col, op, dep, _ = ref.rendering(t_starts, t_ends, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
            sig.grad = None
            rgb.grad = None
            (col.sum() + op.sum() + dep.sum()).backward()

        # ATen's intra-op thread pool does not scale on a problem this small (128 threads: seconds per step): try a
        # few pool sizes on the box's cores and report the fastest, with the size used
        best = None
        for nt in sorted({threads_before, 8, 16, 32, min(64, cores)}):
            if nt > cores:
                continue
            torch.set_num_threads(nt)
            step()
            t = time.perf_counter()
            for _ in range(3):
                step()
            sec = (time.perf_counter() - t) / 3
            if best is None or sec < best[0]:
                best = (sec, nt)
            if sec > 1.0:
                break
        torch.set_num_threads(best[1])
        for _ in range(warmup):
            step()
        t = time.perf_counter()
        for _ in range(iters):
            step()
        sec = (time.perf_counter() - t) / iters
        out = {"impl": "reference pure-PyTorch CPU path (baseline/_ref nerfacc.rendering, batched, config 1)",
               "workload": f"{GRID_RES}^3-scene chords, {n_rays} rays x {n_samples} samples, uniform sigma, fwd+bwd",
               "value": n_rays * n_samples / sec, "unit": "samples/s", "ms_per_step": sec * 1e3,
               "cpu_count": os.cpu_count(), "cores_usable": cores, "torch_num_threads": torch.get_num_threads(),
               "threads_note": "fastest of a few ATen pool sizes on this box", "steps": iters, "warmup": warmup}
        torch.set_num_threads(threads_before)
        return out
    except Exception as ex:
        return {"unavailable": f"{type(ex).__name__}: {ex}"}


def run_cpu_arm(args, rank, world):
    """`--impl reference`: CPU arm.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as orc
    from nerfacc_b200 import scenes
    orc.build()
    orc.set_num_threads(physical_cores(orc))  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core
    n_rays = RAYS_PER_GPU if orc.num_threads() >= 8 else 8192  # whole config-2 batch when the box has the cores
    ro, rd = scenes.ball_rays(n_rays)
    bins, aabbs = scenes.ball_grid(GRID_RES), scenes.nested_aabbs(1)
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_oracle_step(orc, ro, rd, bins, aabbs, scenes.BALL_STEP)
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, t = cpu_oracle_step(orc, ro, rd, bins, aabbs, scenes.BALL_STEP)
        tot_n += n
        tot_t += t
    v = tot_n / tot_t
    sample = f"{n_rays} of the {RAYS_PER_GPU} rays of the same scene per step (oracle port, OpenMP over rays)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(tot_n // max(args.steps, 1) * (RAYS_PER_GPU // n_rays), scenes.BALL_STEP, args.gpus),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": orc.num_threads(), "kind": "port", "sample": sample},
        # config 1 of BASELINE.json: the one CPU implementation the reference itself has (batched rendering)
        "cpu_baseline_torch": time_reference_cpu_torch(),
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_cpu_arm(args, rank, world)
        return

    import torch.distributed as dist
    import nerfacc_b200 as nfa
    from nerfacc_b200 import _lib, parallel, scenes

    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU; there is no CPU fallback"
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()  # long before the timed region: spawning a process next to it skews the ranks
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    pin_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    # ---- synthetic workload: every rank owns a 65 536-ray shard of a world*65 536 batch (weak scaling)
    n_total = RAYS_PER_GPU * world
    ro_all, rd_all = scenes.ball_rays(n_total, seed=42)
    b, e = parallel.shard_bounds(n_total, rank, world)
    ro_h = torch.from_numpy(ro_all[b:e]).pin_memory()
    rd_h = torch.from_numpy(rd_all[b:e]).pin_memory()
    ro_d, rd_d = ro_h.to(dev), rd_h.to(dev)
    R = e - b
    est = nfa.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=GRID_RES, levels=1).to(dev)
    est.binaries = torch.from_numpy(scenes.ball_grid(GRID_RES)).to(dev)
    step_size = scenes.BALL_STEP

    ri, ts, te = est.sampling(ro_d, rd_d, render_step_size=step_size)
    N = ri.numel()
    g = torch.Generator(device="cpu").manual_seed(43 + rank)
    sigmas = (5 * torch.rand(N, generator=g)).to(dev).requires_grad_(True)
    rgbs = torch.rand(N, 3, generator=g).to(dev).requires_grad_(True)
    target = torch.rand(R, 3, generator=g).to(dev)
    loss_host = torch.zeros(1).pin_memory()

    def field(t_starts, t_ends, ray_indices):  # stands in for the user's radiance field
        return rgbs, sigmas

    # Loss reductions in flight.  The reduced loss is consumed LOSS_LAG steps late, as a logger would: reading it
    # makes the compute stream wait for that collective, i.e. for the slowest rank to have reached it, so a lag
    # of one step turns every bit of host jitter on any rank into a stall on all of them.
    pending = []

    # e2e arm: every step's rays come from pinned host memory.  The copy of step k+1 is issued on a copy stream as
    # soon as step k's march has consumed its rays (double buffer), so it overlaps step k's compositing -- what a
    # data loader does; one H2D copy of the full ray batch per step stays inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [(torch.empty_like(ro_d), torch.empty_like(rd_d)) for _ in range(2)]
    staged = {"event": None, "slot": 0}

    def prefetch_rays():
        slot = staged["slot"] ^ 1
        with torch.cuda.stream(copy_stream):
            stage[slot][0].copy_(ro_h, non_blocking=True)
            stage[slot][1].copy_(rd_h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged["event"], staged["slot"] = ev, slot

    def read_back(value):
        loss_host.copy_(value, non_blocking=True)  # the (reduced) loss goes back to the host

    ticket = [None]  # pipelined arm: the traversal of the NEXT step, already queued

    def step(host_inputs, pipelined: bool = False):
        if host_inputs:
            if staged["event"] is None:
                prefetch_rays()  # first step of the arm: nothing was prefetched yet
            torch.cuda.current_stream(dev).wait_event(staged["event"])
            o, d = stage[staged["slot"]]
            # the copy of the NEXT step's rays is started while this step's sampling() waits for its march: the
            # buffer it overwrites was last read by the previous step's march, which is complete by then
            nfa.defer_until_wait(prefetch_rays)
        else:
            o, d = ro_d, rd_d
        if pipelined:
            if ticket[0] is None:
                ticket[0] = est.sampling_begin(o, d, render_step_size=step_size)
            ri_, ts_, te_ = est.sampling_end(ticket[0])
        else:
            ri_, ts_, te_ = est.sampling(o, d, render_step_size=step_size)
        colors, opac, depth, _ = nfa.rendering(ts_, te_, ri_, n_rays=R, rgb_sigma_fn=field)
        loss = torch.nn.functional.mse_loss(colors, target)
        if pipelined:  # the next batch's traversal goes in front of this batch's backward kernels
            ticket[0] = est.sampling_begin(o, d, render_step_size=step_size)
        # the only collective of the path: it overlaps the backward, and its host-side enqueue is parked until the
        # next sampling() waits for its march (the step is host-bound, that wait is the host's only idle time)
        red = parallel.all_reduce_loss_async(loss, defer=LOSS_DEFER and world > 1, transport=LOSS_TRANSPORT) \
            if COLLECTIVE else \
            parallel.LossReduction(loss.detach(), 1.0, reduce=False)
        sigmas.grad = None
        rgbs.grad = None
        with torch.autograd.set_multithreading_enabled(False):  # one GPU per process: skip the engine's thread hop
            loss.backward()
        pending.append(red)
        if len(pending) > LOSS_LAG and not NOWAIT:  # consumed LOSS_LAG steps late, as a logger would
            done = pending.pop(0)
            if host_inputs:
                nfa.defer_until_wait(lambda: read_back(done.result()))  # D2H read in the next march wait
            else:
                done.result()
        return ri_.numel()

    ranks_ms = []  # local ms/step of every rank in the most recent timed() call

    def timed(host_inputs: bool, steps: int, warmup: int, clocks=None, load_steps: int = 0, pipelined: bool = False):
        if clocks:
            clocks.mark_load()
        for _ in range(warmup + load_steps):  # the same count on every rank: steps contain a collective
            step(host_inputs, pipelined)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launches
        e0.record()
        n = 0
        for _ in range(steps):
            n += step(host_inputs, pipelined)
        _lib.run_idle_tasks()  # deferred chores of the last step, then the trailing collectives / read-backs:
        while pending:         # everything completes inside the timed region
            if NOWAIT:         # (debugging aid: the mailbox ring has long been overwritten -- nothing to collect)
                pending.clear()
                break
            done = pending.pop(0).result()
            if host_inputs:
                read_back(done)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ck = clocks.stop() if clocks else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            print(f"[rank {rank}] host_inputs={host_inputs} local ms/step={ms / steps:.4f}", file=sys.stderr, flush=True)
        t = torch.tensor([ms, float(n)], device=dev, dtype=torch.float64)
        ranks_ms[:] = [ms / steps]
        if world > 1:
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            ranks_ms[:] = [float(x[0]) / steps for x in every]  # the limiter of an N-rank step is visible here
            ms, n = max(float(x[0]) for x in every), sum(float(x[1]) for x in every)
        return ms, float(n), _lib.launches - l0, ck

    ms, n_samples, launches, ck = timed(False, args.steps, args.warmup, clocks, CLOCK_LOAD_STEPS)
    per_rank = {"min": round(min(ranks_ms), 4), "max": round(max(ranks_ms), 4), "all": [round(x, 4) for x in ranks_ms]}
    loss_route = "none (1 GPU)"
    if world > 1:
        mb = parallel.PeerMailbox._instances.get((dev.type, dev.index))
        if mb is not None:
            mb.check()
        loss_route = ("NVLink peer mailbox (csrc/peer.cu)" if mb is not None else "NCCL all_reduce") + \
            f", deferred into the march wait, read {LOSS_LAG} step(s) late"
    value = n_samples / (ms * 1e-3)
    # the e2e arm gets the same warm-up as the value arm plus a short untimed run-in: its first steps allocate the
    # staging buffers, open the copy stream and re-size the allocator's pools
    ms_e2e, n_e2e, _, _ = timed(True, args.steps, args.warmup, load_steps=min(CLOCK_LOAD_STEPS, 200))
    # extra, NOT the headline: the same step with sampling_begin / sampling_end (an extension of the drop-in API,
    # see OccGridEstimator.sampling_begin), the next batch's traversal queued before this batch's backward
    ms_pipe, n_pipe, _, _ = timed(False, args.steps, args.warmup, pipelined=True)
    if ticket[0] is not None:
        est.sampling_end(ticket[0])
        ticket[0] = None
    e2e_value = n_e2e / (ms_e2e * 1e-3)

    # ---- per-kernel roofline: CUDA events on the launching stream, L2 flushed before each launch
    roof = None
    cpu_base = None
    ref_cuda = None
    cpu_torch = None
    if rank == 0:
        peak, peak_kind = load_peaks()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        pi = nfa.pack_info(ri, R)
        gcol = torch.rand(R, 3, device=dev)

        def time_call(fn, setup=None, reps=10):
            tot = 0.0
            fn(setup() if setup else None)  # untimed first call (lazy initialisation inside torch)
            for _ in range(reps):
                arg = setup() if setup else None
                flush.fill_(1)
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn(arg)
                c.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(c)
            return tot / reps * 1e-3

        # The kernels are timed as KERNELS: launched straight through the C ABI with the argument lists
        # nerfacc_b200.volrend / grid use, so the events bracket the launch alone and not ~80 us of autograd /
        # Python in front of it (the GPU would sit idle inside the bracket and the host would be billed to HBM).
        f32 = dict(dtype=torch.float32, device=dev)
        sg, cl = sigmas.detach(), rgbs.detach()
        w_o, t_o, a_o = torch.empty(N, **f32), torch.empty(N, **f32), torch.empty(N, **f32)
        c_o, o_o, d_o, raw = (torch.empty((R, 3), **f32), torch.empty((R, 1), **f32), torch.empty((R, 1), **f32),
                              torch.empty((R, 5), **f32))
        g_sg, g_cl = torch.empty(N, **f32), torch.empty((N, 3), **f32)
        P = _lib.ptr

        def k_fwd(_):
            _lib.call("nfa_composite_fwd", dev, R, N, P(pi), P(ts), P(te), P(sg), 0, P(cl), None, None, 1, P(w_o), P(t_o),
                      P(a_o), P(c_o), P(o_o), P(d_o), P(raw))

        def k_bwd(_):  # gradient w.r.t. colours only, as the MSE-on-colours step produces
            _lib.call("nfa_composite_bwd", dev, R, N, P(pi), P(ts), P(te), P(sg), 0, P(cl), None, None, 1, P(raw),
                      P(gcol), None, None, None, None, None, P(g_sg), P(g_cl))

        # the two traversal kernels, launched the way OccGridEstimator.sampling launches them
        from nerfacc_b200.grid import _MarchJob
        job = _MarchJob(ro_d, rd_d, est.binaries, est.aabbs, None, None, step_size, None, None, None,
                        want_intervals=False, want_terminate=False, near_plane=0.0, far_plane=1e10)

        def k_march(_):
            job._launch_march()

        def k_expand(_):  # expand (writes packed_info as well: no separate offsets pass on this route)
            job._expand_samples(N)

        def k_trav(_):
            est.sampling(ro_d, rd_d, render_step_size=step_size)

        t_march, t_expand = time_call(k_march), time_call(k_expand)
        job.sc.busy = False
        job.sc = None
        t_fwd, t_bwd, t_trav = time_call(k_fwd), time_call(k_bwd), time_call(k_trav)
        # what this GPU does on a write-only stream of the expand kernel's size (ATen fill, timed the same way): the
        # copy figure in MEASURED_PEAKS.json is read + write; a pure store stream is slower, and expand is one
        wbuf = torch.empty(B_TRAVERSE * N, dtype=torch.uint8, device=dev)
        t_fill = time_call(lambda _: wbuf.fill_(1))
        del wbuf
        # parity of the direct launches with the public API (same kernels, same arguments)
        with torch.no_grad():
            chk = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b_, c: (cl, sg))
        assert torch.equal(chk[0], c_o) and torch.equal(chk[1], o_o), "direct kernel launch differs from rendering()"
        # algorithmic bytes per launch (SURVEY 8d per-unit figures x units per launch):
        #   march   reads rays 24 B/ray + writes counts 8 B/ray (+ ~32 B per run record, ~1 run/ray)
        #   expand  16 B/sample written + packed_info 16 B/ray
        #   fwd     44 B/sample + 36 B/ray,  bwd 48 B/sample + 40 B/ray
        # and the bytes the kernels must really move (they never read the 8 B/sample of ray_indices: the
        # packed_info stashed by sampling() replaces it) -- `frac_dram` uses those.
        stages = {
            "march": (64 * R, 64 * R, t_march),
            "expand": (B_TRAVERSE * N + 16 * R, B_TRAVERSE * N + 48 * R, t_expand),
            "composite_fwd": (B_FWD * N + 36 * R, (B_FWD - 8) * N + 36 * R, t_fwd),
            "composite_bwd": (B_BWD * N + 40 * R, (B_BWD - 8) * N + 40 * R, t_bwd),
        }
        dom = max(stages, key=lambda k: stages[k][2])  # the kernel that takes longest, whatever it is
        by, by_real, tt = stages[dom]
        traffic = load_traffic(dom, N)
        roof = {"bound": "hbm", "kernel": dom, "achieved": by / tt / 1e9, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": by / tt / 1e9 / peak, "algorithmic_bytes": by,
                "traffic": traffic,
                "frac_algorithmic": by / tt / 1e9 / peak,
                "frac_dram": (traffic if traffic else by_real) / tt / 1e9 / peak,
                "frac_dram_source": "ncu dram bytes (profiles/r2_traffic.json)" if traffic else
                                    "bytes the kernel must move (no ray_indices read)",
                "stages_us": {k: round(v[2] * 1e6, 1) for k, v in stages.items()},
                "stages_frac_algorithmic": {k: round(v[0] / v[2] / 1e9 / peak, 3) for k, v in stages.items()},
                "stages_frac_dram": {k: round((load_traffic(k, N) or v[1]) / v[2] / 1e9 / peak, 3)
                                     for k, v in stages.items()},
                "write_only": {"gbs": round(B_TRAVERSE * N / t_fill / 1e9, 1), "fill_us": round(t_fill * 1e6, 1),
                               "how": "torch fill_ of 16 B x n_samples, timed like the stages",
                               "expand_frac": round(t_fill / t_expand, 3)},
                "sampling_call_us": round(t_trav * 1e6, 1),
                "step_frac_of_roofline": ((B_TRAVERSE + B_FWD + B_BWD) * N + B_RAY * R) / (ms / args.steps * 1e-3) / 1e9 / peak}
        del flush
        if world == 1 and not args.no_reference_cuda:
            ref_cuda = time_reference_cuda(dev, ro_d, rd_d, step_size, R, N)
        if not args.no_cpu_baseline:
            from oracle import oracle as orc
            orc.build()
            os.sched_setaffinity(0, all_cpus)  # the CPU baselines may use every host core again
            orc.set_num_threads(physical_cores(orc))
            bins, aabbs = scenes.ball_grid(GRID_RES), scenes.nested_aabbs(1)
            ro_c, rd_c = ro_all[b:e], rd_all[b:e]  # the whole 65 536-ray batch, as the reference arm runs it
            cpu_oracle_step(orc, ro_c, rd_c, bins, aabbs, step_size)
            tn, tsec = 0, 0.0
            t_begin = time.perf_counter()
            while tsec < 3.0 and time.perf_counter() - t_begin < 30.0:
                n_, t_ = cpu_oracle_step(orc, ro_c, rd_c, bins, aabbs, step_size)
                tn += n_
                tsec += t_
            cpu_base = {"value": tn / tsec, "unit": "samples/s", "cores": orc.num_threads(), "kind": "port",
                        "sample": f"all {R} rays of the same scene, repeated for ~3 s (oracle port, OpenMP over rays)"}
            cpu_torch = time_reference_cpu_torch()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(N, step_size, world),
            "details": {"loss_all_reduce": loss_route,
                        "l2": "per-step working set ~0.9 GB > 126 MB L2 (inputs larger than L2); per-kernel "
                              "roofline timings flush L2 before each launch",
                        "per_rank_ms_per_step": per_rank},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(R * 24),
                    "d2h_bytes_per_step": 4 + 32, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "pipelined": {"value": n_pipe / (ms_pipe * 1e-3), "unit": "samples/s", "ms_per_step": ms_pipe / args.steps,
                          "note": "extra: sampling_begin/sampling_end, next batch's traversal queued before backward(); "
                                  "not the drop-in API, not the headline"},
            "clocks": ck,
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "reference_cuda": ref_cuda,
            "cpu_baseline_torch": cpu_torch,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        parallel.PeerMailbox.shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
