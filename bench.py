#!/usr/bin/env python
"""Benchmark of the render/refine hot path.  Prints ONE JSON line (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): rays/s, forward + backward, 256x256 render of one
DeepSDF latent.  A *step* is one full refine iteration of the reference's
``Optimizer.optimize`` body (pipelines/optimizer.py:79-157) for ONE detection at
256x256 with the 40^3 lattice (cfg2 of SURVEY.md 8(d)): DeepSDF lattice eval with
input gradient -> band extraction -> surfel splat -> 2D + 3D losses -> every
gradient -> Adam/SGD update.  rays/s = detections * W * H / step time.

``value``  : device-resident inputs, CUDA-event timed, L2 flushed between steps.
``e2e``    : the same step through the public API ``Optimizer.optimize(1, ...)`` with
             host (pinned) inputs: H2D of the NOCS prediction, LIDAR crop, K and
             parameters, D2H of parameters + loss history, every step.
``roofline``: the dominant kernel (DeepSDF MLP forward + input-gradient) timed alone.
``cpu_baseline`` / ``--impl reference``: the oracle restatement of the reference's
             algorithm (torch CPU fp32, all host threads, decoder weights left
             requiring grad so the two wasted dW passes of the reference are paid).
N > 1: frames are sharded, one process per GPU, no data-path collective (weak scaling);
the label all-gather happens after the timed region (dump time).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 256
DENSITY = 40
PRIOR = os.path.join(ROOT, "assets", "deepsdf_synth.pt")
FLOP_PER_POINT_FWD = None   # filled from the decoder spec


def mlp_flops_per_point(spec_json):
    """2 * sum(in*out) per point-evaluation: 3 671 040 for the stock spec."""
    ns = spec_json["NetworkSpecs"]
    L = spec_json["CodeLength"]
    full = [L + 3] + list(ns["dims"]) + [1]
    tot = 0
    for l in range(len(full) - 1):
        out = full[l + 1] - full[0] if (l + 1) in ns.get("latent_in", []) else full[l + 1]
        tot += full[l] * out
    return 2 * tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_scene():
    """Synthetic cfg2 inputs, generated once by oracle/make_bench_scene.py and committed."""
    g = np.load(os.path.join(ROOT, "assets", f"bench_scene_{SIZE}.npz"))
    return {"K": g["K"], "crop_size": [int(v) for v in g["crop_size"]], "density": int(g["density"]),
            "nocs_pred": g["nocs_pred"], "lidar": g["lidar"], "weights": {"2d": float(g["w2d"]), "3d": float(g["w3d"])},
            "init": {k: g["init_" + k] for k in ("yaw", "trans", "scale", "latent")}}


def load_oracle_prior():
    """Only the cpu_baseline / --impl reference legs touch the oracle."""
    from oracle import prior as P
    return P.load_prior(PRIOR)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm
# ------------------------------------------------------------------------------------------------
def cpu_iteration_time(prior, sc, steps, warmup, threads=None):
    import torch
    from oracle import sdf_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    for t in list(prior.weight) + list(prior.bias):   # reference keeps requires_grad=True on the decoder
        t.requires_grad_(True)
    pts = O.lattice(DENSITY)
    st = O.RefineState.create(**sc["init"])
    K = torch.from_numpy(sc["K"])
    nocs = torch.from_numpy(sc["nocs_pred"])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.refine_iteration(prior, pts, K, SIZE, SIZE, st, nocs, sc["lidar"], sc["weights"]["2d"], sc["weights"]["3d"],
                           tile_rows=16)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    for t in list(prior.weight) + list(prior.bias):
        t.requires_grad_(False)
    return float(np.mean(times)), threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    prior, sc = load_oracle_prior(), load_scene()
    steps, warmup = max(1, min(args.steps, 5)), min(args.warmup, 1)
    sec, threads = cpu_iteration_time(prior, sc, steps, warmup)
    value = SIZE * SIZE / sec
    sample = (f"{steps} full refine iteration(s) at {SIZE}x{SIZE}, D={DENSITY}, 1 detection (oracle port of the "
              f"reference algorithm, torch CPU fp32, splat evaluated in 16-row pixel tiles, 2D loss windowed, "
              f"decoder dW passes kept as in the reference), {warmup} warm-up")
    line = {
        "impl": "reference", "metric": "rays/s (fwd+bwd) 256x256 DeepSDF render", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(batch):
    return {"workload": f"cfg2: one refine iteration (fwd+bwd+update) per step, {batch} detection(s)/GPU, "
                        f"{SIZE}x{SIZE} crop, Grid3D({DENSITY}) = {DENSITY**3} lattice points, stock 8x512 DeepSDF "
                        f"prior (synthetic, latent 3), 400 LIDAR points",
            "detections_per_gpu": batch, "width": SIZE, "height": SIZE, "density": DENSITY,
            "l2": "flushed between timed steps (256 MiB write)", "parallelism": "frames sharded, no collective"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    from sdflabel_b200 import _lib
    from sdflabel_b200.deepsdf.workspace import setup_dsdf
    from sdflabel_b200.grid import Grid3D
    from sdflabel_b200.pipelines.optimizer import Optimizer, _engine_for

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION (some images export it) prints a banner there
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    sc = load_scene()
    spec_json = json.load(open(os.path.splitext(PRIOR)[0] + ".json"))
    flop_pt = mlp_flops_per_point(spec_json)
    B = args.batch

    dec, L = setup_dsdf(PRIOR, precision=torch.float32)
    dec = dec.to(dev)
    if args.mlp == "ffma":
        dec.mlp_impl = _lib.MLP_FFMA
    elif args.mlp == "tcgen05":
        dec.mlp_impl = _lib.MLP_TCGEN05
    grid = Grid3D(DENSITY, device=dev)
    K = torch.from_numpy(sc["K"])
    nocs = torch.from_numpy(sc["nocs_pred"]).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # ---- device-resident loop: engine iterations ------------------------------------------
    eng = _engine_for(dec, B, DENSITY, SIZE, SIZE, sc["lidar"].shape[0], 64, sc["weights"], dec.mlp_impl)
    for b in range(B):
        eng.set_detection(b, K, SIZE, SIZE, nocs, sc["lidar"], sc["init"]["yaw"], sc["init"]["trans"],
                          sc["init"]["scale"], sc["init"]["latent"])
    sampler = ClockSampler(local_rank)     # samples clocks / throttle reasons over every timed section below
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        eng.run(1)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches0 = lib.sdfr_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for s0, s1 in evs:
        flush.fill_(1)                      # L2 flush, outside the timed events
        s0.record()
        eng.run(1)
        s1.record()
    torch.cuda.synchronize()
    launches = lib.sdfr_launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(np.sum(step_ms))
    if dist:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
    ms_per_step = total_ms / args.steps
    value = world * B * SIZE * SIZE / (ms_per_step * 1e-3)
    params_after, hist = eng.get(0)

    # ---- end-to-end through the public API with host buffers ------------------------------------
    e2e_times = []
    h2d = nocs.numel() * 4 + sc["lidar"].size * 4 + (1 + 3 + 1 + L) * 4 + 2 * 9 * 4
    d2h = (5 + L) * 4 + 4 * 4
    params = {k: v.copy() for k, v in sc["init"].items()}
    opt = Optimizer(params, dev, sc["weights"])
    for i in range(3 + args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"], viz_type=None)   # ends with the D2H read
        torch.cuda.synchronize()
        if i >= 3:
            e2e_times.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_times))
    if dist:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * SIZE * SIZE / e2e_s

    # ---- roofline of the dominant kernel: the DeepSDF MLP forward over the lattice ---------------
    # (the engine evaluates the input gradient only for the ~2.5 % of the lattice inside the band, so the
    # forward-only lattice launch is the dominant kernel of a step)
    ng = DENSITY ** 3
    lat = torch.nn.functional.normalize(torch.from_numpy(sc["init"]["latent"]), dim=0).to(dev).repeat(B, 1).contiguous()
    sdf = torch.empty(B * ng, device=dev)
    dinp = torch.empty(B * ng, L + 3, device=dev)
    impl = dec.mlp_impl if dec.mlp_impl else (_lib.MLP_TCGEN05 if dec.native().tcgen05 else _lib.MLP_FFMA)
    # the engine's lattice pass: fp16-operand (hi-only) forward when the tensor-core decoder is in use
    k_impl = _lib.MLP_TCGEN05_COARSE if impl == _lib.MLP_TCGEN05 else impl
    kev = []
    for i in range(3 + args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.sdfr_decoder_eval_lattice(dec.native().handle, lat.data_ptr(), B, DENSITY, sdf.data_ptr(),
                                                 0, k_impl, _lib.stream_ptr()))
        b.record()
        if i >= 3:
            kev.append((a, b))
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    alg_flops = 1.0 * flop_pt * ng * B          # forward over the lattice (F flop per point)
    achieved = alg_flops / (k_ms * 1e-3) / 1e12
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["bf16_tflops"], "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    else:
        peak, peak_src = 1590.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "mlp_dram_bytes.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get("bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": _lattice_kernel_name(impl),
                "kernel_ms": k_ms, "share_of_step": k_ms / ms_per_step, "peak_source": peak_src,
                "algorithmic_flops_per_launch": alg_flops,
                "issued_over_algorithmic": 1.0,
                "note": "lattice pass of the engine (forward, one fp16 MMA per product); the accurate 3-MMA "
                        "forward+gradient pass runs on the pre-selected band points only"}

    clocks = sampler.stop() if rank == 0 else None

    # ---- dump-time exchange (N > 1): all-gather of the label records ---------------------------------
    if dist:
        rec = torch.tensor(np.concatenate([[rank], params_after]).astype(np.float32), device=dev)
        out = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(out, rec)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            sec, threads = cpu_iteration_time(load_oracle_prior(), sc, 1, 0)
            cpu = {"value": SIZE * SIZE / sec, "unit": "rays/s", "cores": threads, "kind": "port",
                   "sample": f"1 full refine iteration at {SIZE}x{SIZE}, D={DENSITY} (oracle port, torch CPU fp32, "
                             f"16-row pixel tiles, decoder dW passes kept as in the reference), {sec:.1f} s"}
        line = {
            "metric": "rays/s (fwd+bwd) 256x256 DeepSDF render", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3, "api": "sdflabel_b200.pipelines.optimizer.Optimizer.optimize(1, ...)"},
            "roofline": roofline, "cpu_baseline": cpu,
            "frames_per_s": world * B / (ms_per_step * 1e-3) / 60.0,
            "notes": {"frames_per_s": "detections/s assuming the reference's 60 iterations per detection",
                      "final_loss": float(hist[-1, 2]) if len(hist) else None,
                      "mlp_impl": "tcgen05" if impl == _lib.MLP_TCGEN05 else "ffma"},
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def _lattice_kernel_name(impl):
    """Name of the kernel the lattice pass launches (the dispatch of csrc/mlp_tc.cu:launch_mlp_tc_coarse)."""
    from sdflabel_b200 import _lib
    if impl != _lib.MLP_TCGEN05:
        return "mlp_ffma_kernel"
    if os.environ.get("SDFR_TC_PINGPONG", "1") == "0":
        return "mlp_tc_kernel"
    if os.environ.get("SDFR_TC_PAIR", "1") != "0":
        return "mlp_tc_coarse_pair_kernel"
    return "mlp_tc_coarse_wide_kernel" if os.environ.get("SDFR_TC_WIDE", "1") != "0" else "mlp_tc_coarse_kernel"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="detections per GPU per step")
    ap.add_argument("--mlp", default="auto", choices=["auto", "ffma", "tcgen05"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
