#!/usr/bin/env python
"""bench.py -- superposed samples/s of the B200-native SuperDiff sampler (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the CPU path of the reference, oracle port)

A bench "step" is ONE full sampling call: T reverse-diffusion steps, 2 UNets per step, batch 64 per
GPU at 256x256 (BASELINE.json configs[2] on the reference UNet architecture).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "superposed samples/sec at 256^2 (2 UNets, 250 steps)"
UNIT = "samples/s"
CONV_MAC_PER_PIXEL = 664713          # all 10 convs of one UNet forward
MAC_128x128 = 147456                 # one 128->128 3x3 conv, per pixel


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the two roofline kernels from the committed
    `ncu --set full` captures (profiles/r1_final_ncu.md); taken at exactly the launch shapes timed below."""
    p = os.path.join(ROOT, "profiles", "r1_final_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    conv = next((v for k, v in d.items() if k.startswith("conv3x3")), None)
    upd = next((v for k, v in d.items() if k.startswith("superpose_update")), None)
    return (conv["traffic_MB"] * 1e6 if conv else None), (upd["traffic_MB"] * 1e6 if upd else None)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.index, self.samples, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if self.mask & b], "samples": len(s)}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_reference_run(R, T, M, sample_B, n_steps, warm_steps=1):
    """Time the oracle (the reference's UNet restated + our A7 oracle) on the host cores for a bounded
    sample: sample_B images, n_steps of the T diffusion steps; returns (samples/s extrapolated, seconds)."""
    from oracle import superdiff_oracle as O
    torch.set_num_threads(os.cpu_count())
    params = [O.init_unet_params(i) for i in range(M)]
    sched = O.Schedule(T)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((sample_B, 1, R, R), generator=g)
    logq = torch.zeros(sample_B, M)
    dt = 0.0
    with torch.no_grad():
        for k in range(warm_steps + n_steps):
            t = T - 1 - k
            z = torch.randn(x.shape, generator=g)
            t0 = time.perf_counter()
            tt = torch.full((sample_B,), t, dtype=torch.long)
            eps = [O.unet_forward(p, x, tt) for p in params]
            x, logq, _ = O.superpose_step(x, eps, z, logq, sched.alphas[t], sched.alpha_bars[t], sched.betas[t])
            if k >= warm_steps:
                dt += time.perf_counter() - t0
    per_step = dt / n_steps
    return sample_B / (per_step * T), dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    R, T, M = args.res, args.diffusion_steps, 2
    sB, sN = 4, 4
    for _ in range(args.warmup):
        cpu_reference_run(R, T, M, sB, 1, warm_steps=0)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        v, _ = cpu_reference_run(R, T, M, sB, sN, warm_steps=0)
        vals.append(v)
    wall = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    sample = f"B={sB}, {sN} of {T} diffusion steps per bench step, extrapolated linearly in T"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    which = {(128, 100): "BASELINE configs[1]", (256, 250): "BASELINE configs[2]",
             (512, 1000): "BASELINE configs[3] per-GPU shard"}.get((args.res, args.diffusion_steps), "custom shape")
    return {"workload": f"TB+Pneumonia superposition {args.res}x{args.res}, batch {args.batch} per GPU, "
                        f"{args.diffusion_steps}-step DDPM schedule ({which}, reference UNet architecture)",
            "per_gpu_batch": args.batch, "global_batch": args.batch * world, "resolution": args.res,
            "diffusion_steps": args.diffusion_steps, "models": 2, "parallelism": f"batch-shard x{world}",
            "l2": "working set per call >> L2 (no flush needed between calls)",
            "noise": "in-kernel Philox (value) / host noise stack (e2e)"}


# ----------------------------------------------------------------------------- dominant-kernel roofline leg
def conv_roofline(S, dev, R, chunk, iters=20, cin=128, cout=128, impl=2, flush_l2=True):
    """Mean duration of the dominant kernel (tcgen05 128->128 3x3 conv) at the shape the sampler launches it
    with (chunk samples of RxR): CUDA events around each launch on the launching stream (sdd_conv3x3_profile),
    256 MiB rewritten before every launch to flush L2."""
    import ctypes
    lib = S.lib()
    act = torch.randn(chunk, R, R, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.03
    bias = torch.zeros(cout, device=dev)
    out = torch.empty(chunk, R, R, cout, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = ctypes.c_float()
    rc = lib.sdd_conv3x3_profile(act.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), chunk, R, R, cin, cout,
                                 impl, iters, flush.data_ptr(), flush.numel() if flush_l2 else 0, ctypes.byref(ms),
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.sdd_last_error()
    flops = 2.0 * 9 * cin * cout * chunk * R * R
    return flops / (ms.value * 1e-3) / 1e12, ms.value


def update_roofline(S, dev, B, D, iters=20, noise=False, rotating=True):
    """Fused superposition-update kernel alone: M=2, fp32 eps; in-kernel Philox => 16 B/element
    (explicit noise tensor => 20 B/element).  rotating: `iters` back-to-back launches between one CUDA-event pair,
    each on the next of several buffer sets totalling >= 512 MB (4 x L2), so every launch works on data that is not
    in L2 while the kernel's code stays warm; otherwise: one event pair per launch with an L2 flush before each
    (which also evicts the instructions: +8..10 us on a 15 us kernel)."""
    import ctypes
    lib = S.lib()
    ms = ctypes.c_float()
    if rotating:
        rc = lib.sdd_superpose_update_profile_rotating(B, D, 2, 1 if noise else 0, iters, 512 << 20, ctypes.byref(ms),
                                                       torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.sdd_last_error()
    else:
        x = torch.randn(B, D, device=dev)
        eps = torch.randn(2, B, D, device=dev)
        z = torch.randn(B, D, device=dev) if noise else None
        logq = torch.zeros(B, 2, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        rc = lib.sdd_superpose_update_profile(x.data_ptr(), eps.data_ptr(), z.data_ptr() if noise else None,
                                              logq.data_ptr(), B, D, 2, iters, flush.data_ptr(), flush.numel(),
                                              ctypes.byref(ms), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.sdd_last_error()
    bpe = 20.0 if noise else 16.0
    return bpe * B * D / (ms.value * 1e-3) / 1e9, ms.value


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (weak scaling)")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--diffusion-steps", type=int, default=250)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the kernel-alone roofline legs (ncu launch lists)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    import super_diff_disease_b200 as S

    if not torch.cuda.is_available():
        raise S.SddError("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, R, T, M = args.batch, args.res, args.diffusion_steps, 2
    D = R * R
    models = []
    for i in range(M):
        torch.manual_seed(i)  # random-init weights, PyTorch default init of the reference architecture
        models.append(S.UNet().to(dev))
    ddpm = S.DDPM(T)
    lo = rank * B  # weak scaling: every rank samples its own B images, global ids [rank*B, (rank+1)*B)
    shape = (B, 1, R, R)
    gathered = [torch.empty(shape, device=dev) for _ in range(world)] if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_call(seed):
        x = S.superposed_sample(models, ddpm, shape, dev, seed=seed, sample_offset=lo)
        if world > 1:
            dist.all_gather(gathered, x)  # the path's only collective: final sample gather
        return x

    # ---- roofline legs: the two roofline kernels timed ALONE (burst-peak denominators), before the long timed region
    # heats the board into its power cap (measured after it, the same launches read 4-8 % lower and vary run to run)
    roof = None
    if rank == 0 and not args.no_roofline:
        chunk = int(os.environ.get("SDD_CHUNK", "0")) or max(1, min(B, (1536 << 20) // (R * R * 128 * 2)))
        conv_roofline(S, dev, R, chunk, iters=3)  # warm-up: module load, attributes, clocks
        conv_tf, conv_ms = conv_roofline(S, dev, R, chunk)
        update_roofline(S, dev, B, D, iters=50)
        upd_gbs, upd_ms = update_roofline(S, dev, B, D, iters=200)
        upd_gbs_n, upd_ms_n = update_roofline(S, dev, B, D, iters=200, noise=True)
        roof = (chunk, conv_tf, conv_ms, upd_gbs, upd_ms, upd_gbs_n, upd_ms_n)
        torch.cuda.empty_cache()
    barrier()

    for w in range(args.warmup):
        one_call(1000 + w)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        one_call(2000 + k)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    _, launches = S.superposed_sample(models, ddpm, shape, dev, seed=1, sample_offset=lo, return_launches=True)
    torch.cuda.synchronize()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = B * world * args.steps / (ms * 1e-3)

    # ---- e2e: public API with HOST buffers (pinned noise stack in, samples out), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        S.sampling.clear_cache()
        stack_h = torch.empty((T, B, 1, R, R), dtype=torch.float32).pin_memory()
        stack_h.normal_(generator=torch.Generator().manual_seed(rank))
        out_h = torch.empty(shape, dtype=torch.float32).pin_memory()
        stack_d = torch.empty_like(stack_h, device=dev)

        def e2e_call():
            stack_d.copy_(stack_h, non_blocking=True)
            x = S.superposed_sample(models, ddpm, shape, dev, noise=stack_d)
            if world > 1:
                dist.all_gather(gathered, x)
            out_h.copy_(x, non_blocking=True)

        e2e_call()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(1, min(args.steps, 2))
        f0.record()
        for _ in range(n_e2e):
            e2e_call()
        f1.record()
        barrier()
        ems = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = t.item()
        e2e = {"value": B * world * n_e2e / (ems * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(stack_h.numel() * 4), "d2h_bytes_per_step": int(out_h.numel() * 4),
               "calls": n_e2e, "note": "noise stack [T,B,1,H,W] from pinned host memory, samples read back to host"}
        del stack_d, stack_h

    if rank == 0:
        hbm, tf_burst, tf_sust, src = peaks()
        if roof is None:  # --no-roofline (launch-list runs): the line carries no roofline claim
            nan = float("nan")
            roof = (max(1, min(B, (1536 << 20) // (R * R * 128 * 2))), nan, nan, nan, nan, nan, nan)
        chunk, conv_tf, conv_ms, upd_gbs, upd_ms, upd_gbs_n, upd_ms_n = roof
        conv_traffic, upd_traffic = ncu_traffic()
        if not (B == 64 and R == 256 and chunk == 64):
            conv_traffic = upd_traffic = None  # the captures were taken at the default launch shapes only
        step_flops = 2.0 * CONV_MAC_PER_PIXEL * D * M * T  # per sample
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, world), "clocks": clk,
                "gpu_launches": int(launches) * args.steps * world,
                "whole_path_tensor_frac_of_sustained": value / world * step_flops / (tf_sust * 1e12),
                "roofline": {"kernel": "conv3x3_tc4_kernel<128> (GN+SiLU+conv 128->128, 66.6% of conv FLOPs)", "bound": "tensor",
                             "achieved": conv_tf, "peak": tf_burst, "unit": "TFLOP/s", "frac": conv_tf / tf_burst,
                             "traffic": conv_traffic, "algorithmic_bytes": 2.0 * chunk * R * R * 128 * 2, "peak_source": f"{src} bf16 burst", "launch_ms": conv_ms,
                             "how": f"kernel alone at the sampler's launch shape ({chunk}x{R}x{R}x128), CUDA events "
                                    "around each launch on the launching stream, L2 flushed between launches, taken "
                                    "before the long timed region (board not yet power-capped)"},
                "roofline_update": {"kernel": "superpose_update_kernel<2>", "bound": "hbm", "achieved": upd_gbs,
                                    "peak": hbm, "unit": "GB/s", "frac": upd_gbs / hbm, "traffic": upd_traffic,
                                    "bytes_per_element": 16, "launch_ms": upd_ms, "peak_source": f"{src} copy",
                                    "noise_tensor_variant": {"bytes_per_element": 20, "achieved": upd_gbs_n,
                                                             "frac": upd_gbs_n / hbm, "launch_ms": upd_ms_n},
                                    "how": "kernel alone, M=2, in-kernel Philox noise; 200 back-to-back launches "
                                           "between one CUDA-event pair on the launching stream, each launch on the "
                                           "next of several buffer sets totalling >= 512 MB (inputs larger than L2, "
                                           "code stays warm)"},
                }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            t0 = time.perf_counter()
            v, secs = cpu_reference_run(R, T, M, 4, 8)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"B=4, 8 of {T} diffusion steps after 1 warm-up step "
                                              f"({secs:.1f} s of CPU work), extrapolated linearly in T"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
