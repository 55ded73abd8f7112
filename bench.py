#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: MobileViT-S images/sec at batch 256 (256x256) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config s256|xs_sweep|s512|gru] [--weak]

A "step" is one forward pass of the hot path (mobilevit_model::extract_features, main.cpp:604-646, generalised to a
batch) over one batch of synthetic images.  One process per GPU (torchrun for N>1).  BASELINE configs[2] as written:
batch 256 in TOTAL, sharded 256/128/64/32 per GPU on 1/2/4/8 GPUs (strong scaling, the default; `--weak` keeps 256 per
GPU).  The batch shards as independent per-GPU sub-batches with no data-path collective; every rank writes the logits of
its images into its slice of ONE shared host buffer (the "final host-side gather", ggml_experiments_b200/shard.py) and
rank 0 checks that buffer bit for bit against its own single-GPU run of the whole batch.  torch.distributed is used
only for barriers and the max-over-ranks of the timings.

  value         device-resident: inputs already in HBM, outputs stay in HBM; CUDA events on the launching stream
  e2e           through the host API: H2D of the step's raw u8 images + device-side preprocessing + forward + D2H of
                features and logits + host gather, every step (f32-image variant reported next to it)
  roofline      dominant kernel of the step: algorithmic bytes (SURVEY 8d minimum: `bytes_min`) / its CUDA-event time /
                measured peak; `frac_moved` is the same with the bytes this plan actually moves
  cpu_baseline  the CPU oracle (a port of the reference's ggml algorithm; upstream ggml itself is not available) on a
                bounded sample of the same workload, all host cores

Other BASELINE configs, same schema (not the driver's default line):
  --config xs_sweep   configs[1]: MobileViT-XS, batch 1..256 on one GPU (value = batch 256, `sweep` has every batch)
  --config s512       configs[3]: MobileViT-S at 512x512, 64 images per GPU
  --config gru        configs[4]: rnn_text_gen GRU cell over 4096 streams, tokens/s
`--impl reference` times the oracle port alone (rank 0 only), same metric/config.
"""
from __future__ import annotations

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
sys.path.insert(0, ROOT)

GFLOP_PER_IMAGE = {"s": 4.000, "xs": 2.051, "xxs": 0.814}  # at 256x256, SURVEY.md 8(d)
ACT_MB_PER_IMAGE = {"s": 73.1, "xs": 57.8, "xxs": 25.3}     # layer-wise f16 activation traffic, SURVEY.md 8(d)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops_burst": float(p["bf16_tflops"]),
                "tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(weight_path: str, variant: str, hw: int, n_images: int, threads: int):
    """images/s of the CPU oracle port on a bounded sample (test infrastructure used as the reported baseline)."""
    from ggml_experiments_b200 import weights as W
    from oracle import binding
    m = binding.OracleModel(weight_path)
    imgs = W.synthetic_images(n_images, hw, hw, seed=7)
    _, _, secs = m.forward(imgs, 0, threads, return_time=True)
    return n_images / secs, secs


def metric_name(args):
    if args.config == "gru":
        return f"rnn_text_gen GRU tokens/sec over {args.streams} independent streams"
    total = args.global_images
    return f"MobileViT-{args.variant.upper()} images/sec at batch {total}" + (" in total" if args.world > 1 and not args.weak else "")


def workload_config(args):
    if args.config == "gru":
        return {"workload": f"rnn_text_gen GRU cell (embed 256, units 1024, vocab 66) batched over {args.streams} independent streams, "
                            f"{args.gru_steps} greedy steps per timed step, random-init weights in the gru.bin layout",
                "streams": args.streams, "tokens_per_step": args.streams * args.gru_steps, "parallelism": "1 GPU"}
    in_mb = args.batch * args.hw * args.hw * 12 / 1e6
    return {"workload": f"MobileViT-{args.variant.upper()} forward (extract_features), conv weights f16, {args.hw}x{args.hw} synthetic images, "
                        f"random-init weights in convert-tf-to-ggml layout",
            "variant": args.variant, "per_gpu_batch": args.batch, "global_batch": args.global_images, "image": args.hw,
            "parallelism": f"independent sub-batches x{args.world} (no collective; host-side gather of the logits into one shared buffer)",
            "l2": ("inputs larger than L2: %.0f MB of f32 images per step per GPU" if in_mb > 126 else
                   "%.0f MB of f32 images per step per GPU (< L2), evicted between steps by the forward itself, which streams its whole activation arena through L2") % in_mb}


def run_reference(args, weight_path: str):
    """--impl reference: the CPU oracle port alone, all host threads, a bounded sample of the workload per step."""
    from oracle import binding
    binding.build()
    cores = binding.lib().mvo_max_threads()
    if args.config == "gru":
        import numpy as np
        from oracle import gru_oracle as GO
        w = GO.make_synthetic_gru(seed=5)
        streams, steps = 256, 10  # bounded sample: the numpy restatement is single-threaded BLAS-free matmul per step
        first = (np.arange(streams) * 7 % 66).astype(np.int32)
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            GO.generate_batch(w, first, steps)
            times.append(time.perf_counter() - t0)
        value = streams * steps * args.steps / sum(times)
        desc = f"{streams} of {args.streams} streams x {steps} of {args.gru_steps} steps per timed step, numpy restatement of gru_forward (rnn.cpp:186-263)"
        unit = "tokens/s"
    else:
        sample = min(args.batch, max(8, cores))
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_oracle_rate(weight_path, args.variant, args.hw, min(sample, cores), cores)
        times = []
        for _ in range(args.steps):
            _, secs = cpu_oracle_rate(weight_path, args.variant, args.hw, sample, cores)
            times.append(secs)
        value = sample * args.steps / sum(times)
        desc = f"{sample} of {args.global_images} images per step, batch-1 graphs looped over {cores} host threads"
        unit = "images/s"
    tot = sum(times)
    cfg = dict(workload_config(args), reference_sample=desc)
    return {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f16 conv operands / f32 accumulate, f32 dense (ggml CPU semantics)",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": desc,
                         "note": "CPU oracle restatement of the reference's ggml algorithm; upstream ggml is not vendored (oracle/_ref runs the "
                                 "reference's own programs on the same restatement, single-threaded like main.cpp:640)"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def ctypes_void(x):
    import ctypes
    return ctypes.c_void_p(x)


def run_gru(args, rank, world):
    """BASELINE configs[4]: the batched GRU generator through the ggml boundary, loop resident on the device."""
    import numpy as np
    import torch
    from ggml_experiments_b200 import mobilevit as MV
    from ggml_experiments_b200.gru import GRU
    from oracle import gru_oracle as GO
    assert torch.cuda.is_available()
    MV.set_mode(MV.FAST)
    w = GO.make_synthetic_gru(seed=5)
    path = os.path.join(tempfile.mkdtemp(prefix="gru_bench_"), "gru.bin")
    GO.write_gru_bin(path, w)
    m = GRU(path)
    B, T = args.streams, args.gru_steps
    first = (np.arange(B) * 7 % 66).astype(np.int32)
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        m.generate(first, T)
    n_before = len(sampler.rows)
    dev_ms, wall = 0.0, time.perf_counter()
    for _ in range(args.steps):
        toks, state, ms = m.generate(first, T)  # ms: CUDA-event time of the T-step device loop
        dev_ms += ms
    wall = time.perf_counter() - wall
    sampler.rows = sampler.rows[n_before:]
    clocks = sampler.stop()
    ref, margins, _ = GO.generate_batch(w, first[:64], 16)
    agree = float((toks[:16, :64] == ref).mean())
    peaks = load_peaks()
    flop = 2.0 * B * (1024 * 3072 + 1024 * 66)  # per step; the embedding projection is a folded table lookup
    value = B * T * args.steps / (dev_ms / 1e3)
    out = {"metric": metric_name(args), "value": value, "unit": "tokens/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f16 operands / f32 accumulate (tcgen05), f32 gates and state", "data": "synthetic", "config": workload_config(args),
           "e2e": {"value": B * T * args.steps / wall, "unit": "tokens/s", "h2d_bytes_per_step": int(first.nbytes), "d2h_bytes_per_step": int(toks.nbytes + state.nbytes),
                   "api": "gru_generate (include/gru_b200.h): first tokens up, T device-resident steps, all chosen tokens + final state down"},
           "gpu_launches": 9 * T * args.steps, "clocks": clocks,
           "roofline": {"kernel": "gemm_tcgen05 (recurrent 4096x3072x1024 + dense)", "bound": "tensor", "achieved": round(flop * T * args.steps / (dev_ms / 1e3) / 1e12, 2),
                        "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": round(flop * T * args.steps / (dev_ms / 1e3) / 1e12 / peaks["tflops_sustained"], 4),
                        "traffic": None, "note": "whole step (GEMMs + gate kernel + argmax + feedback copies) against the tensor peak; strictly sequential over T"},
           "cpu_baseline": None, "token_agreement_with_numpy_oracle_first_16_steps": agree}
    m.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="s256", choices=["s256", "xs_sweep", "s512", "gru"])
    ap.add_argument("--variant", default=None, choices=["s", "xs", "xxs"])
    ap.add_argument("--batch", type=int, default=None, help="images in TOTAL (default 256; with --weak: images per GPU)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --batch images per GPU (the default is BASELINE configs[2] as written: "
                                                        "--batch images in total, batch/N per GPU)")
    ap.add_argument("--strong", action="store_true", help="(default since round 2; kept for compatibility)")
    ap.add_argument("--hw", type=int, default=None)
    ap.add_argument("--mode", default=os.environ.get("GGML_B200_MODE", "fast"), choices=["fast", "exact"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4096, help="--config gru: independent streams")
    ap.add_argument("--gru-steps", type=int, default=200, help="--config gru: greedy steps per timed step")
    ap.add_argument("--allow-no-cuda-graph", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.world = world
    defaults = {"s256": ("s", 256, 256), "xs_sweep": ("xs", 256, 256), "s512": ("s", 512, 64), "gru": ("s", 256, 256)}[args.config]
    args.variant = args.variant or defaults[0]
    args.hw = args.hw or defaults[1]
    if args.config == "s512":
        args.weak = True  # configs[3] is specified per GPU: 64 images per GPU
    total = args.batch or defaults[2]
    if args.weak:
        args.batch, args.global_images = total, total * world
    else:
        if total % world:
            raise SystemExit(f"batch {total} is not divisible by {world} GPUs")
        args.batch, args.global_images = total // world, total

    from ggml_experiments_b200 import weights as W
    tmpdir = tempfile.mkdtemp(prefix=f"mvit_bench_r{rank}_")
    weight_path = os.path.join(tmpdir, "weight.ggml")

    if args.impl == "reference":
        if rank != 0:
            return 0
        W.write_weight_file(weight_path, W.make_synthetic_weights(args.variant, seed=1234))
        print(json.dumps(run_reference(args, weight_path)), flush=True)
        return 0

    if args.config == "gru":
        if rank == 0:
            print(json.dumps(run_gru(args, rank, world)), flush=True)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    import ggml_experiments_b200 as G
    from ggml_experiments_b200 import mobilevit as MV
    from ggml_experiments_b200 import shard

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback in the product path)"
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = G.lib_ggml()
    L.ggml_b200_set_device(local_rank)
    stream = torch.cuda.current_stream()
    L.ggml_b200_set_stream(ctypes_void(stream.cuda_stream))
    MV.set_mode(MV.FAST if args.mode == "fast" else MV.EXACT)

    W.write_weight_file(weight_path, W.make_synthetic_weights(args.variant, seed=1234))
    model = G.MobileViT(weight_path)
    n, h, w = args.batch, args.hw, args.hw
    model.prepare(n, h, w)

    # synthetic inputs: ONE global batch (image g = pattern g % 16; image 0 = the reference's own test pattern, main.cpp:680-688);
    # this rank owns the contiguous slice [lo, hi) of it
    lo, hi = shard.shard_range(args.global_images, rank, world)
    assert hi - lo == n
    base = W.synthetic_images(16, h, w, seed=7)
    host_in = model.host_input(n, h, w)  # written in place into the library's pinned input buffer (main.cpp:627-634 flow)
    for i in range(n):
        host_in[i] = base[(lo + i) % 16]
    feat, pooled = model.compute(n, h, w)  # first full pass: H2D + forward + D2H (also warms everything up)
    pooled0 = pooled.copy()
    assert np.isfinite(feat).all()
    gather = shard.HostGather(f"mvit_bench_gather_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}",
                              args.global_images, pooled.shape[1], rank, world, dist if world > 1 else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_rate(mdl, nb, steps, warm):
        """device-resident: CUDA-graph replays back to back, CUDA events on the launching stream, max over ranks"""
        for _ in range(warm):
            mdl.forward_device(nb, h, w)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            mdl.forward_device(nb, h, w)
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- device-resident throughput ("value") -------------------------------------------------------------
    # the clock sampler (nvidia-smi -lms 100) needs a few hundred ms to deliver its first row, more on an 8-GPU box, while the
    # timed region is ~0.1 s: it is started before the warm-up so that it is already streaming during the timed steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        model.forward_device(n, h, w)
    barrier()
    n_before = len(sampler.rows)
    dev_ms = device_rate(model, n, args.steps, 0)
    info = model.plan_info(n, h, w)  # re-read after the replays: `cuda_graph` says whether they went through the captured graph
    extra_s = 0.0
    if len(sampler.rows) - n_before < 2:  # region shorter than the sampling period: keep the identical load running until sampled
        t_extra = time.perf_counter()
        while len(sampler.rows) - n_before < 2 and time.perf_counter() - t_extra < 3.0:
            for _ in range(4):
                model.forward_device(n, h, w)
            torch.cuda.synchronize()
        extra_s = time.perf_counter() - t_extra
    sampler.rows = sampler.rows[n_before:]
    clocks = sampler.stop()
    clocks["window"] = "timed steps" if extra_s == 0.0 else f"timed steps + {extra_s:.2f} s of the identical load (region shorter than the 100 ms sampling period)"
    value = args.batch * world * args.steps / (dev_ms / 1e3)
    if info["mode"] == 0 and not info["cuda_graph"] and not args.allow_no_cuda_graph:
        raise SystemExit("bench.py: the plan did not run as a captured CUDA graph (see the library's message above); a launch-bound number "
                         "would be reported.  Pass --allow-no-cuda-graph to measure anyway.")

    sweep = None
    if args.config == "xs_sweep" and world == 1:  # configs[1]: batch sweep 1..256 on one GPU, device-resident
        sweep = []
        for b in (1, 2, 4, 8, 16, 32, 64, 128, 256):
            model.prepare(b, h, w)
            reps = max(args.steps, min(200, 4000 // b))
            ms = device_rate(model, b, reps, 3)
            sweep.append({"batch": b, "images_per_s": round(b * reps / (ms / 1e3), 1), "ms_per_step": round(ms / reps, 4)})
            if b != n:
                model.release(b, h, w)

    # ---- end to end through the host API ("e2e") ------------------------------------------------------------
    # Two pipelined slots (mvit_slot_*): every step uploads ITS batch from pinned host memory, runs the forward,
    # downloads features + logits and writes the logits into this rank's slice of the shared gather buffer; the copies
    # of one slot overlap the kernels of the other.  The synchronous single-call latency (mvit_compute) is reported too.
    for _ in range(2):
        model.compute(n, h, w)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f, p = model.compute(n, h, w)
    torch.cuda.synchronize()
    sync_s = max_over_ranks(time.perf_counter() - t0)
    assert np.array_equal(p, pooled0), "replayed forward is not deterministic"
    for s in range(2):
        model.slot_input(n, h, w, s)[:] = host_in
        model.slot_submit(n, h, w, s)
    for s in range(2):
        fs, ps = model.slot_wait(n, h, w, s)
        assert np.array_equal(ps, pooled0), "pipelined slot result differs from the synchronous call"
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        s = i & 1
        if i >= 2:
            f, p = model.slot_wait(n, h, w, s)    # step i-2 of this slot: its D2H has landed, buffers are reusable
            gather.write(p)
        model.slot_submit(n, h, w, s)      # H2D(step i) + forward + D2H, asynchronous
    for s in range(2):
        f, p = model.slot_wait(n, h, w, s)
        gather.write(p)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = args.batch * world * args.steps / e2e_s
    h2d = n * h * w * 3 * 4
    d2h = int(f.nbytes + p.nbytes)

    # Same pipeline fed with uint8 images: the resize / 1/255 of sam_image_preprocess (main.cpp:538-601) runs on
    # the device, 4x fewer bytes cross PCIe.
    img_u8 = np.clip(np.rint(host_in * 255.0), 0, 255).astype(np.uint8)  # the same pictures as the f32 path, as 8-bit pixels
    for s in range(2):
        model.slot_input_u8(n, h, w, s, h, w)[:] = img_u8
        model.slot_submit_u8(n, h, w, s, h, w)
    for s in range(2):
        fu, pu = model.slot_wait(n, h, w, s)
        assert (pu.argmax(1) == pooled0.argmax(1)).mean() > 0.99, "u8 path disagrees with the f32 path on top-1"
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        s = i & 1
        if i >= 2:
            fu, pu = model.slot_wait(n, h, w, s)
            gather.write(pu)
        model.slot_submit_u8(n, h, w, s, h, w)
    for s in range(2):
        fu, pu = model.slot_wait(n, h, w, s)
        gather.write(pu)
    u8_s = max_over_ranks(time.perf_counter() - t0)
    model.host_input_u8(n, h, w, h, w)[:] = img_u8
    model.compute_u8(n, h, w, h, w)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.compute_u8(n, h, w, h, w)
    u8_sync_s = max_over_ranks(time.perf_counter() - t0)
    e2e_u8 = {"value": args.batch * world * args.steps / u8_s, "unit": "images/s", "h2d_bytes_per_step": int(img_u8.nbytes),
              "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * u8_s / args.steps, "synchronous_call_ms": 1e3 * u8_sync_s / args.steps,
              "api": "mvit_slot_submit_u8 (device-side sam_image_preprocess: the stem reads the quantised u8 images; 256x256 sources need no resize pass), 2 slots in flight, logits gathered into one shared host buffer"}

    # ---- the gathered logits of all ranks == one GPU running the whole batch, bit for bit ---------------------
    barrier()
    gather_check = None
    if rank == 0:
        got = gather.full.copy()  # u8 pipeline, last step of every rank
        glob_u8 = np.clip(np.rint(base[np.arange(args.global_images) % 16] * 255.0), 0, 255).astype(np.uint8)
        g = args.global_images
        if world == 1 or g <= 512:  # one plan for the WHOLE batch on this GPU: what a 1-GPU run of the same job computes
            model.host_input_u8(g, h, w, h, w)[:] = glob_u8
            _, ref = model.compute_u8(g, h, w, h, w)
            how = f"rank 0 ran all {g} images as one batch"
        else:  # weak scaling with a huge global batch: rank 0 re-runs every rank's sub-batch
            ref = np.empty_like(got)
            for r in range(world):
                a, b = shard.shard_range(g, r, world)
                model.host_input_u8(n, h, w, h, w)[:] = glob_u8[a:b]
                ref[a:b] = model.compute_u8(n, h, w, h, w)[1]
            how = f"rank 0 re-ran the {world} sub-batches of {n} images"
        equal = bool(np.array_equal(got, ref))
        gather_check = {"rows": int(g), "width": int(got.shape[1]), "bytes_per_step_per_rank": int(pu.nbytes), "bit_equal_to_single_gpu": equal, "how": how}
        assert equal, "the gathered logits of the N-rank run differ from the single-GPU run of the same images"

    # ---- per-kernel roofline (rank 0) -----------------------------------------------------------------------
    peaks = load_peaks()
    roofline, kernels = None, []
    if rank == 0:
        prof = model.profile(n, h, w, reps=3)
        agg = {}
        for r in prof:
            a = agg.setdefault(r["kernel"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "bytes_min": 0.0, "launches": 0})
            a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["bytes"] += r["bytes"]; a["bytes_min"] += r.get("bytes_min", r["bytes"]); a["launches"] += 1
        tot_ms = sum(a["ms"] for a in agg.values()) or 1.0
        ridge = peaks["tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            ai = a["flops"] / a["bytes_min"] if a["bytes_min"] else 0.0
            bound = "tensor" if ai > ridge else "hbm"
            secs = a["ms"] * 1e-3
            ach = (a["flops"] / secs / 1e12) if bound == "tensor" else (a["bytes_min"] / secs / 1e9)
            peak = peaks["tflops_sustained"] if bound == "tensor" else peaks["hbm_gbs"]
            kernels.append({"kernel": k, "launches": a["launches"], "ms_per_step": round(a["ms"], 4), "share": round(a["ms"] / tot_ms, 4),
                            "bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                            "frac": round(ach / peak, 4), "gflop": round(a["flops"] / 1e9, 3), "mbytes_min": round(a["bytes_min"] / 1e6, 2),
                            "mbytes_moved": round(a["bytes"] / 1e6, 2),
                            "frac_moved": round((a["flops"] / secs / 1e12 if bound == "tensor" else a["bytes"] / secs / 1e9) / peak, 4)})
        if kernels:
            d = kernels[0]
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as tf:
                    traffic = json.load(tf).get(d["kernel"])
            except Exception:
                pass
            roofline = {"kernel": d["kernel"], "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                        "frac": d["frac"], "bytes_min": d["mbytes_min"] * 1e6, "bytes_moved": d["mbytes_moved"] * 1e6, "frac_moved": d["frac_moved"],
                        "traffic": traffic, "traffic_source": "static: mean dram__bytes per launch of this kernel from the committed ncu capture "
                                                              "(profiles/traffic.json), batch 256 per GPU; not measured in this run",
                        "share_of_step": d["share"], "launches_per_step": d["launches"],
                        "peak_source": peaks["source"] + (" sustained" if d["bound"] == "tensor" else " copy bandwidth")}

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import binding
        binding.build()
        cores = binding.lib().mvo_max_threads()
        per_img_s = {"s": 1.1, "xs": 0.6, "xxs": 0.3}[args.variant] * (args.hw / 256.0) ** 2
        sample = args.cpu_sample or int(min(args.batch, max(cores, min(4 * cores, 15.0 * cores / per_img_s))))
        rate, secs = cpu_oracle_rate(weight_path, args.variant, args.hw, sample, cores)
        rate1, secs1 = cpu_oracle_rate(weight_path, args.variant, args.hw, 2, 1)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{sample} of {args.batch} images, batch-1 graphs looped over {cores} threads, {secs:.1f} s",
               "single_thread_images_per_s": rate1,
               "note": "CPU oracle restatement of the reference's ggml algorithm (upstream ggml is not vendored); the reference itself runs 1 thread (main.cpp:640)"}

    if rank == 0:
        gf = GFLOP_PER_IMAGE.get(args.variant, 0.0) * (args.hw / 256.0) ** 2
        per_gpu = value / world
        out = {
            "metric": metric_name(args), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (ggml conv rounding points), f32 residual stream",
            "data": "synthetic", "config": dict(workload_config(args), mode=("fast" if info["mode"] == 0 else "exact")),
            # headline e2e = the reference's own entry: raw u8 images in (what stbi_load hands to sam_image_preprocess, main.cpp:
            # 517-601), features + logits out; the resize / 1/255 runs on the device.  The f32-image variant (the caller has
            # already run sam_image_preprocess on the CPU, i.e. extract_features(sam_image_f32&), main.cpp:604) is reported next to it.
            "e2e": dict(e2e_u8, f32_input={"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                           "ms_per_step": 1e3 * e2e_s / args.steps, "api": "mvit_slot_submit/mvit_slot_wait, 2 slots in flight",
                                           "synchronous_call_ms": 1e3 * sync_s / args.steps,
                                           "synchronous_call_images_per_s": args.batch * world * args.steps / sync_s}),
            "gather": gather_check,
            "gpu_launches": info["launches"] * args.steps,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "model_roofline": {"gflop_per_image": gf, "tflops_achieved_per_gpu": per_gpu * gf / 1e3,
                               "frac_of_tensor_peak": per_gpu * gf / 1e3 / peaks["tflops_sustained"],
                               "layerwise_hbm_bound_images_per_s": peaks["hbm_gbs"] * 1e3 / ACT_MB_PER_IMAGE.get(args.variant, 1e9) / (args.hw / 256.0) ** 2,
                               "frac_of_layerwise_hbm_bound": per_gpu / (peaks["hbm_gbs"] * 1e3 / ACT_MB_PER_IMAGE.get(args.variant, 1e9) / (args.hw / 256.0) ** 2),
                               "peaks": peaks},
            "kernels": kernels,
            "plan": info,
        }
        if sweep is not None:
            out["sweep"] = sweep
        print(json.dumps(out), flush=True)
    gather.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
