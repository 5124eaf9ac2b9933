#!/usr/bin/env python
"""bench.py -- GQ quantize tokens/s at the 2^16 x 16 codebook (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port

A "step" is one eval-mode pass of the drop-in GaussianQuantRegularizer over one batch of
synthetic latents: BASELINE.json configs[1], z[64, 32, 32, 32] (batch 64 at 256 px -> 65 536
tokens, one group of 16 per token) against the 65 536 x 16 iid-Gaussian codebook:
prologue (clamp/exp/log/randn, torch), fused tcgen05 search + exact re-score, dequant.

  value : tokens/s with the latents resident in HBM (module forward), CUDA events, max over ranks
  e2e   : the same call fed from pinned host memory, H2D of z and D2H of the indices inside
          the timed region
  roofline : the search kernel's algorithmic flops 4*g*M*N over its own CUDA-event time
             (events recorded inside the library around that one launch), against
             MEASURED_PEAKS.json bf16 dense
  cpu_baseline : oracle/gq_oracle.c (OpenMP C port of the reference torch path) on the host
             cores, on a bounded row sample of the same workload

Multi-GPU (torchrun, one rank per GPU): rows shard by batch, every rank quantizes its own
64-image batch (weak scaling); the path has no exchange step (the reference runs pure replicas,
eval.py:97-107), so no collective sits in the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

N_SAMPLES = 65536
GROUP = 16
BATCH = 64
HW = 32
CHANNELS = 16
METRIC = "gq_quantize_tokens_per_s"
UNIT = "tokens/s"
WORKLOAD = ("configs[1]: sd3unet_gq_0.25 regularizer pass (quantize + dequant), batch 64 @256px "
            "-> 65536 tokens x 65536x16 codebook")


def synth_latents(batch, seed):
    """D-trained latents of SURVEY.md 8(d): mu ~ N(0, 0.75), logvar ~ N(-1.386, 0.3^2)."""
    import torch
    gen = torch.Generator().manual_seed(seed)
    mu = torch.randn(batch, CHANNELS, HW, HW, generator=gen) * 0.866
    lv = -1.386 + 0.3 * torch.randn(batch, CHANNELS, HW, HW, generator=gen)
    return torch.cat([mu, lv], 1)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": float(p["bf16_tflops"]), "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    The sampler process is started before warm-up (it needs a few hundred ms to come up) and runs
    until the end; `window(t0, t1)` then keeps the samples whose timestamps fall inside the timed
    region (widened by half a second when the region is shorter than the sampling period)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        rows = []
        if self.p is None:
            return rows
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[2]), float(parts[3]), float(parts[4]),
                             [nm for nm, v in zip(self.NAMES, parts[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        return rows

    @staticmethod
    def window(rows, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sel = [r for r in rows if t0 <= r[0] <= t1]
        if len(sel) < 3:
            sel = [r for r in rows if t0 - 0.5 <= r[0] <= t1 + 0.5]
        if not sel:
            return out
        reasons = sorted({x for r in sel for x in r[4]})
        out.update({"sm_mhz": statistics.median(r[1] for r in sel), "sm_max_mhz": sel[-1][2], "reasons": reasons,
                    "samples": len(sel), "power_w_max": max(r[3] for r in sel)})
        return out


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference torch path (the Python reference itself cannot
    travel to the GPU box), all host threads, a bounded row sample of the same workload per step."""
    if rank != 0:
        return
    import numpy as np
    import torch
    import oracle
    from gq_b200_host import host_codebook
    oracle.set_num_threads(0)      # every host core (torchrun pins OMP_NUM_THREADS=1 for its workers)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cb, nlp = host_codebook(N_SAMPLES, GROUP)
    rows = args.ref_rows
    z = synth_latents(max(1, (rows + HW * HW - 1) // (HW * HW)), 1000)
    zz = z.permute(0, 2, 3, 1).reshape(-1, 2 * CHANNELS)[:rows]
    mu, lv = zz[:, :CHANNELS].contiguous(), zz[:, CHANNELS:].contiguous()
    std = torch.exp(0.5 * torch.clamp(lv, -30.0, 20.0))
    lstd = std.log()
    a = (mu.numpy(), std.numpy(), lstd.numpy(), cb, nlp)
    for _ in range(max(1, min(args.warmup, 2))):
        oracle.quantize(*a, 1.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.quantize(*a, 1.0)
    dt = (time.perf_counter() - t0) / args.steps
    value = rows / dt
    sample = f"{rows} rows of the {BATCH * HW * HW}-row batch per step (rows are independent; rate extrapolates linearly)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_samples": N_SAMPLES, "group": GROUP, "sample_rows": rows,
                   "note": "CPU port of the reference torch path (oracle/gq_oracle.c, OpenMP); the python reference "
                           "is absent on the GPU box"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-rows", type=int, default=8192, help="rows per step of the CPU arm")
    ap.add_argument("--cpu-rows", type=int, default=8192, help="rows of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--batch", type=int, default=BATCH)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gq_b200
    from gq_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    L = _lib.lib()
    reg = gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=N_SAMPLES, group=GROUP, backend="cuda").to(dev).eval()
    B = args.batch
    M = B * HW * HW
    # input pool larger than the 126 MB L2: rotate so no step finds its latents cached
    z_bytes = B * 2 * CHANNELS * HW * HW * 4
    pool_n = max(2, int(160e6 // z_bytes) + 1)
    pool = [synth_latents(B, 17 * rank + i).to(dev) for i in range(pool_n)]
    host_pool = [synth_latents(B, 9000 + 17 * rank + i).pin_memory() for i in range(4)]
    torch.cuda.synchronize()

    def step(z):
        zhat, info = reg(z)
        return zhat, info["indices"]

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- warm-up (also builds the prepared codebook and the workspace) ----
    for i in range(max(args.warmup, 3)):
        step(pool[i % pool_n])
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs ----
    # The dominant kernel is bracketed by CUDA events on every 8th step of this same region (an event
    # between two kernels of the chain switches their programmatic-launch overlap off, so bracketing
    # every launch would slow down the thing being measured); the region itself is timed by ONE event pair.
    PROF_EVERY = 8
    L.gq_profile_enable(PROF_EVERY)
    launches0 = gq_b200.launch_count()
    ev_start, ev_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    t_epoch0 = time.time()
    ev_start.record()
    for k in range(args.steps):
        step(pool[k % pool_n])
    ev_stop.record()
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    t_epoch1 = time.time()
    launches = gq_b200.launch_count() - launches0
    L.gq_profile_enable(0)
    import ctypes
    tot = ctypes.c_double(0.0)
    cnt = ctypes.c_int(0)
    L.gq_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
    # the steps run back to back on one stream: first-start -> last-stop is the honest span
    total_ms = ev_start.elapsed_time(ev_stop)
    kern_ms = tot.value / max(cnt.value, 1)

    # ---- e2e: host buffers in, indices out, copies inside the timed region ----
    # Every step copies its 8 MB latent batch from pinned host memory and reads its int64 indices
    # back.  The copies run on their own stream with rotating device input buffers so that the H2D
    # of step k+1 overlaps the kernels of step k, and the D2H of step k's indices (its own stream,
    # ordered after the step by an event) overlaps the kernels of step k+1 -- what a serving loop
    # does; all of it is timed, and the timed region ends only when the last indices are on the host.
    copy_stream = torch.cuda.Stream(device=dev)
    out_stream = torch.cuda.Stream(device=dev)
    comp_stream = torch.cuda.current_stream(dev)
    NBUF = 4    # device input buffers / host index buffers in rotation: the H2D copies may run up to three steps ahead
    z_dev = [torch.empty_like(pool[0]) for _ in range(NBUF)]
    idx_host = [torch.empty(B, CHANNELS // GROUP, HW, HW, dtype=torch.int64).pin_memory() for _ in range(NBUF)]
    ev_in = [torch.cuda.Event() for _ in range(NBUF)]
    ev_free = [torch.cuda.Event() for _ in range(NBUF)]
    ev_out = [torch.cuda.Event() for _ in range(NBUF)]

    def e2e_loop(n):
        for k in range(n):
            sl = k % NBUF
            with torch.cuda.stream(copy_stream):
                if k >= NBUF:
                    copy_stream.wait_event(ev_free[sl])      # step k-NBUF has consumed this input buffer
                z_dev[sl].copy_(host_pool[k % 4], non_blocking=True)
                ev_in[sl].record(copy_stream)
            comp_stream.wait_event(ev_in[sl])
            _, ind = step(z_dev[sl])
            ev_free[sl].record(comp_stream)
            ev_out[sl].record(comp_stream)
            with torch.cuda.stream(out_stream):
                out_stream.wait_event(ev_out[sl])
                idx_host[sl].copy_(ind, non_blocking=True)
                ind.record_stream(out_stream)

    e2e_loop(2 * NBUF)
    torch.cuda.synchronize()
    e_steps = args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    e0.record()
    e2e_loop(e_steps)
    copy_stream.synchronize()
    comp_stream.wait_stream(out_stream)   # the closing event is recorded after the last D2H has landed
    e1.record()
    torch.cuda.synchronize()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    clocks = ClockSampler.window(sampler.stop(), t_epoch0, t_epoch1)

    # ---- max over ranks ----
    t = torch.tensor([total_ms, e2e_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms_max = [float(x) for x in t.tolist()]

    ms_per_step = total_ms / args.steps
    value = world * M / (ms_per_step * 1e-3)
    e2e_value = world * M / (e2e_ms / e_steps * 1e-3)

    pk = peaks()
    flops = 4.0 * GROUP * M * N_SAMPLES          # 64 flop per (row, codeword) pair
    achieved = flops / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"] if total_ms > 2000 else pk["bf16_tflops"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "search_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 search / f32 exact re-score", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_samples": N_SAMPLES, "group": GROUP, "rows_per_gpu": M,
                   "global_batch": B * world, "latents": "D-trained (SURVEY 8d)", "parallelism": f"rows sharded by batch x{world}",
                   "l2": f"inputs rotate over a {pool_n}-batch pool ({pool_n * z_bytes / 1e6:.0f} MB > 126 MB L2); "
                         "the 4 MB codebook image is L2-resident by design"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": z_bytes, "d2h_bytes_per_step": idx_host[0].numel() * 8,
                "note": "H2D on a copy stream, four input buffers in rotation; D2H of the indices on its own stream behind an event; the timed region ends after the last D2H"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "quantize_tc_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic, "kernel_ms": kern_ms, "kernel_launches_timed": cnt.value,
                     "algorithmic_flops_per_launch": flops, "peak_source": pk["source"] + (" (burst)" if peak == pk["bf16_tflops"] else " (sustained)")},
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"]},
        "wall_s_timed_region": t_wall,
    }

    if rank == 0 and world == 1 and args.cpu_rows > 0:
        import oracle
        from gq_b200_host import host_codebook
        oracle.set_num_threads(0)
        cb, nlp = host_codebook(N_SAMPLES, GROUP)
        rows = args.cpu_rows
        # rows exactly as the module builds them on the device (so both sides see equal std / log std)
        zz = pool[0].permute(0, 2, 3, 1).reshape(-1, 2 * CHANNELS)[:rows]
        mu, lv = zz[:, :CHANNELS].contiguous(), zz[:, CHANNELS:].contiguous()
        std = torch.exp(0.5 * torch.clamp(lv, -30.0, 20.0))
        a = (mu.cpu().numpy(), std.cpu().numpy(), std.log().cpu().numpy(), cb, nlp)
        oracle.quantize(a[0][:256], a[1][:256], a[2][:256], cb, nlp, 1.0)
        t0 = time.perf_counter()
        o_idx, _, _, _ = oracle.quantize(*a, 1.0)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": rows / dt, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                                "sample": f"{rows} rows of pool[0] ({dt:.2f} s wall); rows are independent so the rate extrapolates"}
        # parity of the very tensors that were timed
        _, ind = step(pool[0])
        got = ind.permute(0, 2, 3, 1).reshape(-1)[:rows].cpu().numpy()
        line["parity"] = {"rows_checked": rows, "index_mismatches_vs_oracle": int((got != o_idx).sum())}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
