#!/usr/bin/env python
"""Headline benchmark: training images/sec of the UrsoNet hot path, ResNet-50, bf16, 960x600 (padded to 640x960),
ori_resolution 16, batch 32 per GPU (BASELINE.json configs[1]; at 8 GPUs this is configs[4]'s global batch 256).

  python bench.py --gpus N --steps K --warmup W                 our arm (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference path's CPU restatement (oracle/), rank 0 only

One JSON line on stdout (rank 0).  A "step" = forward + losses + backward + clipped SGD update on one synthetic batch.
`value` is measured with the batch resident in HBM; `e2e` includes the pinned-host -> device copy of every step's
images/labels and the device -> host read of the two losses.  `roofline` is measured live with CUDA events per launch.
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
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
    os.environ.pop("NCCL_DEBUG", None)     # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

TRAIN_GFLOP_PER_IMG = 281.02      # BASELINE.md section 2 (cfg2/cfg5): fwd + dgrad + wgrad, 2 FLOP/MAC


def make_cfg(args):
    from ursonet_b200.config import Config
    cfg = Config()
    cfg.NAME = "bench"
    cfg.BACKBONE = args.backbone
    cfg.BOTTLENECK_WIDTH = 32            # pose_estimator.py:776 default
    cfg.BRANCH_SIZE = 1024               # pose_estimator.py:777
    cfg.NR_DENSE_LAYERS = 1              # pose_estimator.py:820
    cfg.ORI_BINS_PER_DIM = args.ori_resolution
    cfg.REGRESS_ORI = bool(getattr(args, "regress_ori", False))   # --classify_ori is the CLI default (pose_estimator.py:786)
    cfg.REGRESS_LOC = True
    cfg.OPTIMIZER = "SGD"
    cfg.IMAGE_RESIZE_MODE = "pad64"
    cfg.IMAGE_MAX_DIM = args.width
    cfg.IMAGE_MIN_DIM = (args.height + 63) // 64 * 64     # pose_estimator.py:856-860: 600 -> 640
    cfg.IMAGES_PER_GPU = args.batch
    cfg.LEARNING_RATE = 0.001
    cfg.update()
    return cfg


def synth_batch(cfg, B, seed):
    """uint8 RGB frames, loc in the SPEED range, orientation soft labels."""
    import torch
    from ursonet_b200 import labels
    g = torch.Generator().manual_seed(seed)
    H, W = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    loc = torch.stack([torch.rand(B, generator=g) * 4 - 2, torch.rand(B, generator=g) * 4 - 2,
                       torch.rand(B, generator=g) * 35 + 5], 1)
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    q = q * torch.where(q[:, 3:4] < 0, -1.0, 1.0)
    if cfg.REGRESS_ORI:
        return img, loc, q.float()
    enc = labels.OrientationEncoder(cfg.ORI_BINS_PER_DIM, cfg.BETA)
    ori = torch.from_numpy(enc.encode(q.numpy())).float()
    return img, loc, ori


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_step_rate(cfg, steps, warmup, budget_s=25.0):
    """The reference path restated on torch-CPU fp32 (oracle/), all host threads, B=2 sample of the same workload."""
    import torch
    from oracle import ursonet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    img, loc, ori = synth_batch(cfg, B, seed=123)
    p = O.init_weights(cfg, seed=0, dtype=torch.float32)
    batch = (O.mold_image(img, torch.float32), loc, ori)
    state = {}
    for _ in range(warmup):
        p, _ = O.train_step(p, state, batch, cfg, lr=1e-3)
    t0 = time.perf_counter()
    done = 0
    for _ in range(max(steps, 1)):
        p, _ = O.train_step(p, state, batch, cfg, lr=1e-3)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return B * done / dt, dt / done * 1e3, done, B, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--height", type=int, default=600)
    ap.add_argument("--ori_resolution", type=int, default=16)
    ap.add_argument("--regress_ori", action="store_true", help="quaternion regression head (BASELINE configs[2])")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="fixed GLOBAL batch (BASELINE configs[4]: 256): each rank takes global/N images per step as "
                         "micro-batches of --batch with gradient accumulation; reported as strong scaling")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="overlapped two-part gradient all-reduce (N > 1, experimental)")
    ap.add_argument("--reserve-sms", type=int, default=0,
                    help="with --overlap: SMs the second backward segment leaves free for NCCL's all-reduce kernel")
    ap.add_argument("--no-residual-mma", action="store_true", help="A/B: residual / fan-in addends added by the epilogue warps instead of the tensor core")
    ap.add_argument("--stage-split", type=int, default=4, help="A/B: leading convs whose weight operands get their own staging launch")
    ap.add_argument("--pair-l2", action="store_true", help="A/B: co-run the dgrad / wgrad launches that share a large gradient tensor")
    ap.add_argument("--no-tail-split", action="store_true", help="A/B: no N-split of the last partial wave (urso_set_tail_split(0))")
    ap.add_argument("--no-wgrad-halo", action="store_true", help="A/B: Engine W loads one operand atom per filter tap (urso_set_wgrad_halo(0))")
    ap.add_argument("--no-pdl", action="store_true", help="A/B: launch the engines without programmatic dependent launch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default=None, help="write the per-launch CUDA-event table here")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = make_cfg(args)
    workload = (f"{args.backbone} ori_resolution={args.ori_resolution} {int(cfg.IMAGE_SHAPE[0])}x{int(cfg.IMAGE_SHAPE[1])} "
                f"(from {args.width}x{args.height}) batch {args.batch}/GPU, SGD+clipnorm, "
                f"{'regress_ori quaternion' if args.regress_ori else 'classify_ori'}")
    metric = "train images/sec ResNet-50 bf16 @960x600"

    if args.impl == "reference":
        if rank != 0:
            return
        W = min(args.warmup, 1)
        rate, ms, done, B, cores = cpu_reference_step_rate(cfg, min(args.steps, 3), W, budget_s=90.0)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": done, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "note": "TF/Keras cannot be installed (py3.12, no network): reference path "
                       "restated on torch-CPU (oracle/ursonet_oracle.py), bounded sample B=2"},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{done} train steps at B={B} of the same workload"},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from ursonet_b200.engine import Engine
    from ursonet_b200 import lib as _lib
    if args.no_pdl:
        _lib.load().urso_set_pdl(0)
    if args.no_residual_mma:
        _lib.load().urso_set_residual_mma(0)
    if args.no_tail_split:
        _lib.load().urso_set_tail_split(0)
    if args.no_wgrad_halo:
        _lib.load().urso_set_wgrad_halo(0)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_micro = 1
    if args.global_batch:
        per_rank = args.global_batch // world
        assert per_rank * world == args.global_batch and per_rank % min(args.batch, per_rank) == 0, "global batch must split evenly"
        args.batch = min(args.batch, per_rank)
        n_micro = per_rank // args.batch
        workload += f", global batch {args.global_batch} = {world} ranks x {n_micro} micro-batches x {args.batch}"
    eng = Engine(cfg, args.batch, training=True, world_size=world, seed=0,
                 reserve_sms=args.reserve_sms if (world > 1 and args.overlap) else 0, pair_l2=args.pair_l2, stage_split=args.stage_split)
    img, loc, ori = synth_batch(cfg, args.batch, seed=rank)
    h_img, h_loc, h_ori = img.pin_memory(), loc.pin_memory(), ori.pin_memory()
    eng.img_u8.copy_(h_img)
    eng.gt_loc.copy_(h_loc)
    eng.gt_ori.copy_(h_ori)
    allreduce = None
    # measured on 2 x B200: the overlapped schedule is not faster (4005 vs 4015 img/s) -- the persistent conv CTAs leave
    # NCCL's kernel no SM to run on until a launch boundary -- so one blocking all-reduce is the default
    ar_async = (lambda g: dist.all_reduce(g, async_op=True)) if world > 1 and args.overlap else None
    if world > 1 and not args.overlap:
        allreduce = lambda g: dist.all_reduce(g)
    use_graph = not args.no_graph
    lr = cfg.LEARNING_RATE

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def train_step():
        if n_micro == 1:
            eng.train_step(lr, allreduce, use_graph, ar_async)
        else:       # gradient accumulation: n_micro x (fwd + bwd), then ONE all-reduce + update
            for m in range(n_micro):
                eng.accumulate(m, n_micro, use_graph)
            eng.apply_update(lr, allreduce, use_graph, ar_async)

    for _ in range(max(args.warmup, 3)):
        train_step()
    # ---------------- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        train_step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = t_ms.item()
    imgs_per_step = world * args.batch * n_micro
    value = imgs_per_step * args.steps / (t_ms / 1e3)
    # ---------------- timed region 2: end to end through the public step API with host buffers
    losses_host = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n_steps, log):
        """n_steps host-fed steps: H2D of every (micro-)batch from pinned memory, D2H of every step's losses."""
        eng.upload_async(h_img, h_loc, h_ori)               # first batch; every later upload overlaps the previous step
        for i in range(n_steps):
            for m in range(n_micro):
                eng.swap_in()                               # uploaded batch -> the buffers the graphs read (waits for the H2D)
                if i + 1 < n_steps or m + 1 < n_micro:
                    eng.upload_async(h_img, h_loc, h_ori)   # next (micro-)batch's H2D (59 MB from pinned memory) on the copy stream
                if n_micro == 1:
                    eng.train_step(lr, allreduce, use_graph, ar_async)
                else:
                    eng.accumulate(m, n_micro, use_graph)
            if n_micro > 1:
                eng.apply_update(lr, allreduce, use_graph, ar_async)
            losses_host[i & 1].copy_(eng.losses, non_blocking=True)      # D2H of this step's losses (pinned, async)
            loss_ev[i & 1].record()
            if i > 0:                                       # the host reads EVERY step's losses, one step behind the GPU,
                loss_ev[(i - 1) & 1].synchronize()          # so that the launch latency of step i+1 is not exposed
                log.append(losses_host[(i - 1) & 1].tolist())
                stamps.append(time.perf_counter())
        loss_ev[(n_steps - 1) & 1].synchronize()
        log.append(losses_host[(n_steps - 1) & 1].tolist())
        stamps.append(time.perf_counter())

    # warm-up of THIS path (copy stream, staging buffers, first pinned H2D / D2H: one-off driver set-up that measured
    # 0..100 ms, i.e. up to +5 ms per step of a 20-step run, when it was left inside the timed region)
    stamps = []
    e2e_loop(max(args.warmup, 3), [])
    loss_log, stamps = [], []
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_loop(args.steps, loss_log)
    e3.record()
    barrier()
    t2 = torch.tensor([e2.elapsed_time(e3)], device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = imgs_per_step * args.steps / (t2.item() / 1e3)
    h2d = n_micro * (h_img.numel() + 4 * h_loc.numel() + 4 * h_ori.numel())
    loss_vals = loss_log[-1]
    assert len(loss_log) == args.steps

    iv = sorted(1e3 * (b - a) for a, b in zip(stamps[:-1], stamps[1:]))
    step_iv = {"median": round(iv[len(iv) // 2], 3), "max": round(iv[-1], 3)} if iv else None   # diagnostic (host clock)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- live per-launch timing (CUDA events on the launching stream) for the roofline
    prof = eng.profile_ops(train=True, reps=3)
    agg = {}
    for r in prof:
        a = agg.setdefault(r["kind"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
        a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["bytes"] += r["bytes"]; a["n"] += r.get("launches", 1)
    if args.profile_json:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_json)), exist_ok=True)
        json.dump({"per_launch": prof, "by_kind": agg}, open(args.profile_json, "w"), indent=1)
    peak_tf, peak_gbs, peak_src = peaks()
    f_ms = agg.get("conv_fwd", {}).get("ms", 0) + agg.get("conv_dgrad", {}).get("ms", 0)
    f_fl = agg.get("conv_fwd", {}).get("flops", 0) + agg.get("conv_dgrad", {}).get("flops", 0)
    f_n = agg.get("conv_fwd", {}).get("n", 0) + agg.get("conv_dgrad", {}).get("n", 0)
    achieved = f_fl / (f_ms * 1e-3) / 1e12 if f_ms else None
    # DRAM bytes per Engine-F launch from the committed ncu pass over the same step (profiles/r0N_traffic.json, newest)
    traffic, traffic_note = None, "no ncu traffic capture for this configuration"
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_traffic.json")))
    tj = cands[-1] if cands else ""
    if os.path.exists(tj) and args.backbone == "resnet50" and args.width == 960 and args.batch == 32 and not args.regress_ori:
        t = json.load(open(tj))["engines"]["conv_gemm"]
        traffic = t["dram_bytes_per_launch"]
        traffic_note = ("ncu dram__bytes_read+write per conv_gemm launch (avg of %d launches); algorithmic bytes per launch %.4g"
                        % (t["launches"], t["algorithmic_bytes_per_launch"]))
    roofline = {"bound": "tensor", "kernel": "conv_gemm_kernel (Engine F: conv fprop + dgrad, %d launches/step)" % f_n,
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if achieved else None,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "flops_per_launch_avg": f_fl / f_n if f_n else None, "ms_per_launch_avg": f_ms / f_n if f_n else None,
                "by_kind": {k: {"ms": round(v["ms"], 4), "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] and v["flops"] else None),
                                "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] and v["bytes"] else None), "launches": v["n"]}
                            for k, v in agg.items()},
                "step_frac_of_tensor_peak": value / world * TRAIN_GFLOP_PER_IMG / 1e3 / peak_tf
                if args.backbone == "resnet50" and args.width == 960 else None}
    out = {"metric": metric, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True,
           "scaling": "strong" if args.global_batch else "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": workload, "parallelism": f"dp{world}", "global_batch": imgs_per_step,
                      "micro_batches_per_step": n_micro,
                      "l2": "per-step working set (bf16 activations + gradients, >8 GB at batch 32) >> 126 MB L2: no flush needed",
                      "cuda_graphs": use_graph, "weights": "Keras-default random init (--weights none)"},
           "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                   "ms_per_step": t2.item() / args.steps, "warmup_steps": max(args.warmup, 3),
                   "host_step_interval_ms": step_iv},
           "gpu_launches": ((eng.count_launches(True) - 2) * n_micro + 2 + (n_micro if n_micro > 1 else 0)) * args.steps,
           "roofline": roofline, "losses_last_step": loss_vals}
    if not args.no_cpu_baseline and world == 1:
        rate, ms, done, B, cores = cpu_reference_step_rate(cfg, 2, 0, budget_s=20.0)
        out["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                               "sample": f"{done} train step(s) at B={B} of the same workload, torch-CPU fp32 restatement "
                                         "of the reference graph (TF/Keras not installable here)"}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
