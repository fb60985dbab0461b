#!/usr/bin/env python
"""Headline benchmark: AST 256x256 training throughput (images/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B_per_gpu] [--impl ours|reference]

A step is the reference loop body (src/ModelTrainer.py:78-88): zero_grad -> AST forward -> "L1" loss
-> backward -> [gradient all-reduce, overlapped] -> clip_grad_norm_(1.0) + Adam, in train mode with
drop_path_rate=0.1, on synthetic raw/reference pairs (rand*2-1).  Per-GPU batch is
fixed (weak scaling; 16/GPU x 8 GPUs = BASELINE config 4's global batch 128).
Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU path (oracle port,
the reference itself is Python and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))

METRIC = "train images/sec @256x256 (AST)"
UNIT = "images/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


def _synthetic(batch, size, seed=2024):
    import torch
    g = torch.Generator().manual_seed(seed)
    raw = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    ref = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    return raw, ref


# ---------------------------------------------------------------------------------------------
def cpu_reference_step_rate(steps, warmup, size=256, batch=1, budget_s=40.0):
    """The reference's CPU path for this workload (oracle port: plain PyTorch-eager fp32, autograd,
    torch.optim.Adam, clip_grad_norm_) on all host cores.  Returns images/s and the thread count."""
    import torch
    from oracle import ast_init, ast_oracle, losses_oracle
    torch.set_num_threads(os.cpu_count())
    sd = ast_init.ast_state_dict(seed=1234)   # plain torch init: the reference arm never loads the product library
    params = {k: v.clone().requires_grad_() for k, v in sd.items() if v.is_floating_point()}
    full = dict(sd)
    full.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    raw, ref = _synthetic(batch, size)
    times = []
    t_begin = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        opt.zero_grad()
        out = ast_oracle.ast_forward(full, raw, img_size=size)
        loss = losses_oracle.l1(out, ref)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        if i >= warmup:
            times.append(time.time() - t0)
        if time.time() - t_begin > budget_s and len(times) >= 1:
            break
    return batch * len(times) / sum(times), torch.get_num_threads(), len(times)


def gpu_eager_step_rate(dev, batch, size, steps, warmup, allow_tf32):
    """The kernel bar: the reference modules' math in PyTorch-eager on the SAME B200
    (cuBLAS / cuDNN / ATen), fp32 storage, train mode with per-sample DropPath scales, the reference loop
    body ModelTrainer.py:78-88 without its per-step .item()/print.  Returns (images/s, ms/step)."""
    import torch
    from oracle import ast_init, ast_oracle, losses_oracle
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.allow_tf32 = allow_tf32
    try:
        sd = {k: v.to(dev) for k, v in ast_init.ast_state_dict(seed=1234).items()}
        params = {k: v.clone().requires_grad_() for k, v in sd.items() if v.is_floating_point()}
        full = dict(sd)
        full.update(params)
        plist = list(params.values())
        opt = torch.optim.Adam(plist, lr=1e-3)
        raw, ref = _synthetic(batch, size)
        raw, ref = raw.to(dev), ref.to(dev)
        blocks = sorted({k[: k.index("norm2")] for k in sd if "norm2.weight" in k})
        rates = torch.linspace(0, 0.1, 8).tolist()   # AST.py:703-705: enc 0..0.1, bottleneck 0.1, dec reversed
        names = ["encoderlayer_0", "encoderlayer_1", "encoderlayer_2", "encoderlayer_3"]
        prob = {}
        for i, n in enumerate(names):
            for b in range(2):
                prob[f"{n}.blocks.{b}."] = rates[2 * i + b]
        for b in range(2):
            prob[f"conv.blocks.{b}."] = 0.1
        for i, n in enumerate(["decoderlayer_0", "decoderlayer_1", "decoderlayer_2", "decoderlayer_3"]):
            for b in range(2):
                prob[f"{n}.blocks.{b}."] = rates[::-1][2 * i + b]

        def one():
            drop = {}
            for pre in blocks:
                p = prob[pre]
                if p > 0:
                    keep = 1.0 - p
                    m = lambda: torch.empty(batch, device=dev).bernoulli_(keep).div_(keep)
                    drop[pre] = (m() if (pre + "attn.w") in sd else None, m())
            opt.zero_grad(set_to_none=True)
            out = ast_oracle.ast_forward(full, raw, img_size=size, drop_scales=drop)
            loss = losses_oracle.l1(out, ref)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(plist, 1.0)
            opt.step()

        for _ in range(warmup):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return batch / (ms / 1e3), ms
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def run_reference_gpu(args):
    """`--impl reference-gpu`: the PyTorch-eager kernel bar alone (one GPU, rank 0)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    args.batch = args.batch or 16
    res = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        rate, ms = gpu_eager_step_rate(dev, args.batch, args.size, args.steps, max(args.warmup, 3), tf32)
        res[name] = {"value": rate, "unit": UNIT, "ms_per_step": ms}
    line = {"impl": "reference-gpu", "metric": METRIC, "value": res["fp32"]["value"], "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": res["fp32"]["ms_per_step"],
            "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"AST {args.size}x{args.size} train step, batch {args.batch}, PyTorch-eager "
                                   "(cuBLAS/cuDNN/ATen) port of the reference modules on the same B200",
                       "batch_per_gpu": args.batch},
            "gpu_eager": res}
    print(json.dumps(line))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 5)
    warmup = min(args.warmup, 1)
    t0 = time.time()
    rate, threads, done = cpu_reference_step_rate(steps, warmup, budget_s=120.0)
    elapsed = time.time() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warmup, "ms_per_step": 1000.0 / rate, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "AST 256x256 train step (L1, Adam, clip 1.0), batch 1, reference CPU path "
                               "(oracle port, PyTorch-eager fp32)", "batch_per_step": 1},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{done} full train steps at batch 1 ({elapsed:.0f} s wall)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import uwr
    from uwr import ops
    from uwr.train import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # per-GPU batch: 16 on one GPU; N >= 2 GPUs run BASELINE config 4 as written, global batch 128 (64 / 32 / 16 per GPU)
    B, S = (args.batch if args.batch > 0 else (16 if world == 1 else max(1, 128 // world))), args.size

    torch.manual_seed(1234)
    model = uwr.AST(img_size=S).to(dev)
    model.train()
    torch.manual_seed(1000 + rank)  # per-rank DropPath stream (SURVEY.md §8e caveat 4)
    comm = None
    if world > 1 and not args.torch_allreduce:
        from uwr.nccl import Communicator
        comm = Communicator(rank, world)      # raw NCCL: the bucket all-reduces are captured inside the step's graph
    step = TrainStep(model, "L1", lr=1e-3, world_size=world, local_batch=B, comm=comm)

    raw_h, ref_h = _synthetic(B, S, seed=2024 + rank)
    raw_h, ref_h = raw_h.pin_memory(), ref_h.pin_memory()
    raw_d, ref_d = raw_h.to(dev), ref_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # kernel launches of one step, counted on an eager step BEFORE the graph exists (a replay re-issues the same
    # kernels; an eager step next to the graph's private memory pool would double the footprint at batch 64)
    l0 = ops.launch_count()
    step(raw_d, ref_d)
    launches_per_step = ops.launch_count() - l0

    # The step is replayed from ONE CUDA graph (uwr.graph.GraphedTrainStep); with N > 1 the raw NCCL bucket all-reduces
    # are captured inside it (--torch-allreduce: two graphs around eager torch.distributed all-reduces instead).
    use_graph = not args.no_graph
    graphed = None
    if use_graph:
        from uwr.graph import GraphedTrainStep
        graphed = GraphedTrainStep(step, raw_d, ref_d, warmup=max(args.warmup, 3))

    def step_resident():
        if graphed is not None:
            graphed.replay()
        else:
            step(raw_d, ref_d)

    def step_eager():
        step(raw_d, ref_d)

    loss_host = torch.zeros(1).pin_memory()

    def step_e2e():
        # every step: pinned host -> device copy of ITS inputs and a device -> host read of its loss.  With the
        # graphed step the copy of the next step's batch runs on a side stream under the current step
        # (uwr.graph.GraphedTrainStep.prefetch); the loss is read back (and waited for) every step.
        if graphed is not None:
            loss, _ = graphed.step_prefetched()
            graphed.prefetch(raw_h, ref_h)           # next step's batch, overlapped with this step's kernels
        else:
            r = raw_h.to(dev, non_blocking=True)
            t = ref_h.to(dev, non_blocking=True)
            loss, _ = step(r, t)
        loss_host.copy_(loss.view(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the reference reads loss.item() every step

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    if graphed is not None:
        graphed.prefetch(raw_h, ref_h)               # batch of the first e2e step
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: one extra instrumented EAGER step (CUDA events per launch), after the graph
    # and its memory pool are gone ----
    import gc
    graphed = None
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    roofline, table = None, None
    if rank != 0:
        step_eager()  # every rank takes part in the instrumented step's all-reduces
    if rank == 0:
        with ops.KernelProfile() as prof:
            step_eager()
        table = prof.table()
        fam = {}
        for r in table:
            f = fam.setdefault(r["kernel"], [0.0, 0.0, 0.0, 0])
            f[0] += r["ms_total"]; f[1] += r["bytes_per_launch"] * r["launches"]
            f[2] += r["flops_per_launch"] * r["launches"]; f[3] += r["launches"]
        hbm, tf_bf16, src = _peaks()
        # dominant kernel = the family with the largest share of the step; its roofline figure is the
        # family aggregate (sum of algorithmic bytes or flops over its launches / sum of their durations)
        fam_name, fv = max(((k, v) for k, v in fam.items() if v[1] > 0 or v[2] > 0), key=lambda kv: kv[1][0])
        top = next(r for r in table if r["kernel"] == fam_name)
        gbs, tfl = fv[1] / fv[0] / 1e6, fv[2] / fv[0] / 1e9
        ai = fv[2] / max(fv[1], 1.0)
        tf32_peak = tf_bf16 / 2.0  # dense TF32 is half the bf16 rate
        tensor_bound = ai > (tf32_peak * 1e12) / (hbm * 1e9)
        if tensor_bound:
            roofline = {"bound": "tensor", "achieved": tfl, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tfl / tf32_peak}
        else:
            roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}
        total_ms = sum(r["ms_total"] for r in table)
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "r2_roofline_traffic.json")
        if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum of one committed `ncu --set full` capture
            with open(tp) as f:
                ent = json.load(f).get(fam_name)
            if ent:
                traffic = ent["dram_bytes_per_launch"]
                traffic_note = {"shape": ent["shape"], "algorithmic_bytes_per_launch": ent["algorithmic_bytes_per_launch"],
                                "source": ent["source"]}
        roofline.update({"traffic": traffic, "traffic_capture": traffic_note, "kernel": fam_name, "launches_per_step": fv[3],
                         "ms_per_launch": fv[0] / fv[3], "share_of_step": fv[0] / total_ms,
                         "arithmetic_intensity": ai,
                         "top_shape": {"shape": top["shape"], "ms_per_launch": top["ms_avg"], "gbs": top["gbs"],
                                       "tflops": top["tflops"], "launches": top["launches"]},
                         "peak_source": f"MEASURED_PEAKS.json ({src}); tf32 peak = bf16 sustained / 2",
                         # context, not the denominator: the copy peak is a 1:1 read:write figure; a plain float4 stream
                         # kernel on the same GPUs reaches 5.5 TB/s at 1 read : 4 writes (the fp32-C streaming shapes of
                         # this family) and 5.9 TB/s at 1:1 (tools/micro/rw_mix.cu, profiles/r2_micro_benchmarks.txt)
                         "mix_ceiling_note": "memory system at 1:4 read:write 5.5 TB/s, 1:1 5.9 TB/s, read-only 7.0 TB/s "
                                             "(tools/micro/rw_mix.cu); frac is against the copy peak as the contract asks"})
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            with open(args.profile_out, "w") as f:
                json.dump({"batch_per_gpu": B, "size": S, "step_ms": ms / args.steps,
                           "families": {k: {"ms": v[0], "gbs": v[1] / v[0] / 1e6 if v[0] else 0,
                                            "tflops": v[2] / v[0] / 1e9 if v[0] else 0, "launches": v[3]}
                                        for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])},
                           "kernels": table}, f, indent=1)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t0 = time.time()
        rate, threads, done = cpu_reference_step_rate(3, 1, budget_s=30.0)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{done} AST 256x256 train steps at batch 1 on the host "
                                  f"({time.time() - t0:.0f} s wall, oracle port of the reference's CPU path)"}

    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager:
        # the kernel bar: same step, same batch, PyTorch-eager on this GPU (after our own timing, graph pools released)
        torch.cuda.empty_cache()
        try:
            gpu_eager = {}
            for name, tf32 in (("fp32", False), ("tf32", True)):
                rate, ems = gpu_eager_step_rate(dev, B, S, min(args.steps, 10), 3, tf32)
                gpu_eager[name] = {"value": rate, "unit": UNIT, "ms_per_step": ems, "speedup": value / rate}
            gpu_eager["what"] = ("reference modules' math (oracle port) in PyTorch-eager on the same B200: cuBLAS / cuDNN / "
                                 "ATen, fp32 storage; 'fp32' = allow_tf32 off, 'tf32' = allow_tf32 on")
        except RuntimeError as e:   # e.g. out of memory next to our pools: report, do not hide
            gpu_eager = {"error": str(e).splitlines()[0][:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if B == 16 else "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": f"AST {S}x{S} train step: fwd + L1 + bwd + clip_grad_norm(1.0) + Adam, "
                                   f"train mode (drop_path 0.1), batch {B}/GPU, global batch {B * world}",
                       "batch_per_gpu": B, "global_batch": B * world, "image": S,
                       "parallelism": f"dp{world}", "cuda_graph": bool(use_graph), "l2": "working set per step >> 126 MB L2 (no flush needed)",
                       "scaling_note": "N = 1: 16 images; N >= 2: global batch 128 (BASELINE config 4), i.e. 64 / 32 / 16 per GPU",
                       "storage": "fp32 residual stream / weights / gradients, fp16 for the LeFF tensors u and gelu'(v) "
                                  "(10-bit mantissa = what a TF32 operand keeps), TF32 tensor-core products (3xTF32 inside "
                                  "attention), fp32 accumulate"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * raw_h.numel() * 4 * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "gpu_eager": gpu_eager,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0,
                    help="images per GPU per step (default: 16 on one GPU, 128 / N on N >= 2 GPUs = BASELINE config 4)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager-on-GPU comparison leg")
    ap.add_argument("--torch-allreduce", action="store_true",
                    help="N > 1: reduce gradients through torch.distributed work objects (two graphs around eager NCCL) "
                         "instead of the raw NCCL communicator captured in one graph")
    ap.add_argument("--no-graph", action="store_true", help="launch the single-GPU step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile-out", default="", help="write the per-kernel event table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
