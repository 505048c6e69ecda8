#!/usr/bin/env python
"""bench.py -- patch-instances/s through the MIL aggregator at N=50 000 x D=1024 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16x3|bf16x3|fp16|bf16] [--pipeline single|pair]

One "step" = one pass of the hot path (abmil.DAttention eval forward: projection -> gated tanh attention logit ->
softmax over N -> weighted pool -> classifier) over one synthetic bag of N=50 000 x D=1024 fp32 (204.8 MB).
N>1 (torchrun): bag-parallel -- every rank streams its own bags, no data-path collective; value = all ranks' instances
divided by the max-over-ranks device time ("scaling": "weak").
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_INST, D_IN, N_CLASSES = 50000, 1024, 2
WORKLOAD = "abmil.DAttention eval forward, N=50000 x D=1024 fp32, act=relu, C=2"
METRIC = "patch-instances/sec through MIL aggregator at N=50k x D=1024"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def tensor_peaks():
    """(burst, sustained) dense bf16 TFLOP/s: MEASURED_PEAKS.json, else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs.  NVML is initialised BEFORE the timed region (the thread
    signals `ready`): importing / initialising it inside the region held the GIL against the first launches of the loop -- a fixed 0.15 ms (one
    rank) to 0.3 ms (eight ranks initialising NVML at once), i.e. 5-10 % of a 20-step run and the whole of round 1's "scaling loss"."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        self.ready, self.active = threading.Event(), False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)                  # first query pays the lazy set-up
            self.ready.set()
            while not self.stop_flag:
                if self.active:
                    self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.005 if self.active else 0.001)
        except Exception as e:  # NVML missing: report it, do not fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm)}


CONFIG = {"workload": WORKLOAD}            # identical in both arms (the driver compares the dicts)


def cpu_forward_fn():
    """The CPU arm: the reference's OWN abmil.DAttention class (unmodified files under oracle/_ref/, shipped to the box by
    oracle/make_ref.py) when present -- kind "reference"; else the oracle restatement of its forward -- kind "port"."""
    import torch
    import cases
    sd = cases.abmil_state(2021)
    try:
        import _refload
        if _refload.have_reference():
            R = _refload.load_reference()
            m = R.abmil.DAttention(D_IN, N_CLASSES, dropout=0.0, act="relu").eval()
            m.load_state_dict(sd, strict=True)
            return (lambda x: m(x)), torch, "reference", f"reference class modules/abmil.py:DAttention ({_refload.REF_ROOT})"
    except Exception as e:                                   # e.g. torchvision missing on the box: fall back to the port, say so
        print(f"bench.py: reference classes unavailable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    from oracle import mil_oracle as O
    return (lambda x: O.abmil_dattention(sd, x, "relu")), torch, "port", "oracle port of abmil.DAttention.forward"


def time_cpu(n_rep, n_inst):
    fn, torch, kind, what = cpu_forward_fn()
    import cases
    torch.set_num_threads(os.cpu_count())
    x = cases.make_bag(2021, n_inst, D_IN)
    with torch.no_grad():
        fn(x)
        ts = []
        for _ in range(n_rep):
            t0 = time.perf_counter()
            fn(x)
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], kind, what


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (torch CPU, all host threads)."""
    if rank != 0:
        return
    n_rep = max(args.steps, 20)
    med, kind, what = time_cpu(n_rep, N_INST)
    val = N_INST / med
    cb = {"value": val, "unit": "instances/s", "cores": os.cpu_count(), "kind": kind,
          "sample": f"{n_rep} full bags of N={N_INST} (median), {what}, torch CPU fp32, {os.cpu_count()} threads"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "instances/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": CONFIG, "cpu_baseline": cb,
                      "e2e": {"value": val, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def sharded_giant_bag(args, model, dev, rank, world, dist, torch, mhimk):
    """BASELINE config 5 beside the headline (N > 1 only): ONE giant bag N = 200 000 x 1024 instance-sharded along N over the ranks;
    per rank the fused pass on its rows, ONE all-gather of the 2056-byte record its tail wrote, one merge + classifier kernel.
    Strong scaling: ms per bag (max over ranks), the share of it that is not the local fused kernel (exchange + merge + host), and
    on rank 0 the same bag on one GPU for the speed-up."""
    from mhimk import dist as D
    n_giant, steps = 200000, max(10, min(args.steps, 30))
    lo, hi = D.row_slices(n_giant, world)[rank]
    shards = [torch.randn(hi - lo, D_IN, device=dev, generator=torch.Generator(device=dev).manual_seed(99 + rank + 100 * i)) for i in range(3)]

    def step(i):
        return D.sharded_abmil_forward(model, shards[i % 3])[0]

    for i in range(5):
        step(i)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    mhimk.ops.profile_fused(True)
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    n_t, k_tot = mhimk.ops.profile_collect()
    mhimk.ops.profile_fused(False)
    t = torch.tensor([ms, k_tot / max(n_t, 1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kernel_ms = (float(v) for v in t.tolist())
    del shards
    single_ms = None
    if rank == 0:                                            # the same giant bag on ONE GPU (819 MB), for the strong-scaling ratio
        full = torch.randn(1, n_giant, D_IN, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        with torch.no_grad():
            for _ in range(3):
                model(full)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(10):
                model(full)
            s1.record()
            torch.cuda.synchronize()
        single_ms = s0.elapsed_time(s1) / 10
        del full
    dist.barrier()
    return {"workload": f"abmil.DAttention eval forward, ONE giant bag N={n_giant} x D={D_IN}, instance-sharded along N x{world}",
            "scaling": "strong", "ms_per_bag": ms, "instances_per_s": n_giant / (ms * 1e-3), "steps": steps,
            "local_fused_kernel_ms": kernel_ms, "exchange_overhead_us": (ms - kernel_ms) * 1e3,
            "single_gpu_ms_per_bag": single_ms, "speedup_vs_single_gpu": (single_ms / ms) if single_ms else None,
            "collective": "one ncclAllGather of (m, l, P[512]) = 2056 B per rank on the compute stream, then mil_shard_merge_cls_f32; latency-bound",
            "launches_per_rank_per_bag": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["fp16x3", "bf16x3", "fp16", "bf16"])
    ap.add_argument("--pipeline", default="auto", choices=["auto", "single", "pair"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import cases
    import mhimk
    from mhimk.modules import DAttention

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = DAttention(D_IN, N_CLASSES, dropout=0.0, act="relu").to(dev).eval()
    model.load_state_dict({k: v.to(dev) for k, v in cases.abmil_state(2021).items()}, strict=True)
    model.precision = args.precision
    if args.pipeline == "auto":                    # the library's own default
        args.pipeline = "pair"
    os.environ["MHIMK_PIPELINE"] = {"single": "1", "pair": "2"}[args.pipeline]
    # 4 distinct bags (820 MB) visited round-robin: every step streams 205 MB that cannot be in the 126 MB L2
    n_bags = 4
    bags = [torch.randn(1, N_INST, D_IN, device=dev, generator=torch.Generator(device=dev).manual_seed(2021 + 17 * rank + i)) for i in range(n_bags)]

    def step(i):                       # the fused pass + classifier, inputs resident in HBM
        with torch.no_grad():
            return model(bags[i % n_bags])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    # The FIRST NCCL collective of a process finishes its connection set-up lazily: the kernel launched right after it starts ~0.25 ms
    # late (tools/probe_step_gaps.py: first step 458 us after the first barrier, 210 us after any later one, 145-150 us otherwise).  One
    # barrier ahead of the warm-up keeps that one-off out of the timed region (it was most of round 1's 6-7 % "scaling loss" at K = 20).
    sync_all()
    for i in range(args.warmup):
        step(i)
    sampler.ready.wait(20)                      # NVML is up before anything is timed
    sync_all()
    sampler.active = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sync_all()
    total_ms = e0.elapsed_time(e1)

    # the dominant kernel alone: CUDA events recorded inside the C ABI around mil_fused_kernel on the launching stream
    mhimk.ops.profile_fused(True)
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    n_timed, k_total = mhimk.ops.profile_collect()
    mhimk.ops.profile_fused(False)
    kernel_ms = k_total / max(n_timed, 1)

    # end to end through the public module call: pinned host bag -> H2D -> forward -> logits D2H, every step.  Two device buffers:
    # the copy of bag i+1 (copy stream) overlaps the forward of bag i (compute stream); every step still moves its own 204.8 MB
    # and reads its own logits back inside the timed region.
    host = [torch.randn(1, N_INST, D_IN).pin_memory() for _ in range(2)]
    dbufs = [torch.empty(1, N_INST, D_IN, device=dev) for _ in range(2)]
    houts = [torch.empty(1, N_CLASSES).pin_memory() for _ in range(2)]
    e2e_steps = max(3, min(args.steps, 10))
    copy_stream, main_stream = torch.cuda.Stream(device=dev), torch.cuda.current_stream(dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        for b in range(2):
            consumed[b].record(main_stream)
        for i in range(n):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])            # the forward that read this buffer two steps ago is done
                dbufs[b].copy_(host[b], non_blocking=True)
                copied[b].record(copy_stream)
            main_stream.wait_event(copied[b])
            with torch.no_grad():
                houts[b].copy_(model(dbufs[b]), non_blocking=True)
            consumed[b].record(main_stream)

    e2e_loop(2)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(main_stream)
    e2e_loop(e2e_steps)
    e3.record(main_stream)
    sync_all()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_ms = e2.elapsed_time(e3)

    # free the headline buffers, then the instance-sharded giant bag (N > 1 only)
    del bags, host, dbufs
    torch.cuda.empty_cache()
    sharded = sharded_giant_bag(args, model, dev, rank, world, dist, torch, mhimk) if world > 1 else None

    t = torch.tensor([total_ms, e2e_ms, kernel_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms = (float(v) for v in t.tolist())

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = total_ms / args.steps
        value = world * N_INST / (ms_step * 1e-3)
        alg_bytes = N_INST * D_IN * 4                     # X read exactly once; weights (2.4 MB) excluded
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_fused.json")
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get(args.precision)
        # the binding bound of the pass: the two contractions run on the tensor cores, x3 in the hi/lo-split parity arithmetic
        nprod = 3 if args.precision in ("bf16x3", "fp16x3") else 1
        flops = (2.0 * D_IN * 512 + 2.0 * 512 * 128) * N_INST * nprod
        tpeak, tsus, tsrc = tensor_peaks()
        tflops = flops / (kernel_ms * 1e-3) / 1e12
        out = {"metric": METRIC, "value": value, "unit": "instances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "precision": args.precision, "pipeline": args.pipeline,
               "config": CONFIG,
               "detail": {"parallelism": f"bag-parallel x{world}", "l2": "4 distinct 205 MB bags round-robin (> 126 MB L2)",
                          "operand_arithmetic": {"fp16x3": "fp16 hi+lo split, 3 tcgen05 products, fp32 accumulate", "bf16x3": "bf16 hi+lo split, 3 tcgen05 products, fp32 accumulate", "fp16": "single fp16 product",
                                                 "bf16": "single bf16 product"}[args.precision]},
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                            "peak_source": peak_src, "kernel_ms": kernel_ms, "alg_bytes": alg_bytes,
                            "note": "the fused kernel alone (mil_fused2_kernel for the pair pipeline), CUDA events on its stream inside the C ABI, mean over the timed steps"},
               "roofline_tensor": {"bound": "tensor", "achieved": tflops, "peak": tpeak, "peak_sustained": tsus, "unit": "TFLOP/s", "frac": tflops / tpeak,
                                   "peak_source": tsrc, "flops_per_launch": flops,
                                   "note": "issued tensor-core FLOPs of the two contractions (D->512, 512->128) x products per operand pair; the 3-product "
                                           "parity arithmetic is tensor-bound (<= ~29 % of the HBM roofline at the measured peaks), see DESIGN.md 4.1"},
               "e2e": {"value": world * N_INST / (e2e_ms / e2e_steps * 1e-3), "unit": "instances/s", "h2d_bytes_per_step": alg_bytes,
                       "d2h_bytes_per_step": N_CLASSES * 4, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                       "h2d_gbs_per_rank": alg_bytes / (e2e_ms / e2e_steps * 1e-3) / 1e9, "h2d_gbs_aggregate": world * alg_bytes / (e2e_ms / e2e_steps * 1e-3) / 1e9,
                       "note": "PCIe-bound: the pinned-host -> HBM copy of the 204.8 MB bag is the step; all ranks copy concurrently"},
               "gpu_launches": args.steps,                # per step: ONE fused kernel (merge + classifier in its tail; weight images cached)
               "clocks": sampler.summary()}
        if not args.no_cpu_baseline and world == 1:        # rank 0 at N=1 only: at N>1 the other ranks' host threads would distort it
            n_cpu = 250                                    # bounded sample: ~10 s of CPU work on the box's host cores
            med, kind, what = time_cpu(n_cpu, N_INST)
            out["cpu_baseline"] = {"value": N_INST / med, "unit": "instances/s", "cores": os.cpu_count(), "kind": kind,
                                   "sample": f"{n_cpu} full bags of N={N_INST} (median {med * 1e3:.1f} ms), {what}, torch CPU fp32, all host threads"}
        if sharded is not None:
            out["sharded"] = sharded
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
