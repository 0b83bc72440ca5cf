#!/usr/bin/env python
"""bench.py — 10 s-clip waveform -> scores throughput (clips/s) of the B200 path, beside the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips C] [--impl b200|reference]

A "step" is one pass of the hot path (fused log-mel -> VGGish -> multi-level-attention head) over one batch of
C synthetic 10 s / 16 kHz clips per GPU (default 256 = BASELINE.json configs[1]).  For N > 1 the driver launches
one rank per GPU with torchrun; clips are sharded by batch, the path has no collective, so the scaling
is weak (C clips per GPU) and NCCL is used only for the barrier and the max-over-ranks of the timings.

One JSON line is printed by rank 0:
  value      clips/s over all GPUs, inputs resident in HBM (device-timed with CUDA events, max over ranks)
  e2e        the same metric through the host-buffer C-ABI call (vmb_pipeline_forward_host): pinned host input,
             H2D copy, compute, D2H copy of the scores inside the timed region
  roofline   the dominant kernel (tcgen05 implicit GEMM: 5 convs + 3 FCs per step) against the measured bf16 peak
  cpu_baseline  the oracle (numpy front end + torch-CPU VGGish + head = the reference's algorithm) on the host cores
--impl reference times that CPU path alone (rank 0 only) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU measurement on rank 0 alone and must
    # use every host core it can (set before numpy / torch load their thread pools)
    _ncpu = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_k] = _ncpu

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "10s-clip audio->logits clips/sec"
UNIT = "clips/s"
MODEL_CONF = (2, 1)
N_CLASSES = 527
CLIP_SAMPLES = 160000

# algorithmic FLOPs per 0.96 s example: 2*M*N*K per GEMM
_CONV = ((96 * 64, 64, 9), (48 * 32, 128, 576), (24 * 16, 256, 1152), (24 * 16, 256, 2304), (12 * 8, 512, 2304),
         (12 * 8, 512, 4608))
_FC = ((12288, 4096), (4096, 4096), (4096, 128))
FLOP_CONV = [2.0 * m * n * k for m, n, k in _CONV]
FLOP_FC = [2.0 * a * b for a, b in _FC]
FLOP_IGEMM_PER_EXAMPLE = sum(FLOP_CONV[1:]) + sum(FLOP_FC)          # what the tcgen05 kernel computes
FLOP_DFT_PER_CLIP = 2.0 * 400 * 514 * 998
FLOP_MEL_PER_CLIP = 2.0 * 257 * 64 * 998
FLOP_HEAD_PER_CLIP = 29.69e6
FLOP_PER_CLIP = FLOP_DFT_PER_CLIP + FLOP_MEL_PER_CLIP + 10 * (sum(FLOP_CONV) + sum(FLOP_FC)) + FLOP_HEAD_PER_CLIP
STAGES = ("logmel", "conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "fc1", "fc2", "fc3", "mla",
          "postprocess")



def _json_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when
    the communicator is created), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved
    original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


_OUT = None


def emit(line):
    (_OUT or sys.stdout).write(json.dumps(line) + "\n")
    (_OUT or sys.stdout).flush()

def igemm_traffic():
    """DRAM bytes per launch of the dominant kernel, from the committed ncu capture (None if absent)."""
    try:
        with open(os.path.join(ROOT, "profiles", "igemm_traffic.json")) as fh:
            return float(json.load(fh)["dram_bytes_per_launch"])
    except Exception:
        return None


def tensor_pipe_util():
    """Tensor-pipe utilisation per layer from the committed ncu capture (BASELINE.json's second metric); a profiler
    figure, copied from profiles/, never measured inside the timed run."""
    try:
        with open(os.path.join(ROOT, "profiles", "igemm_traffic.json")) as fh:
            return json.load(fh)["tensor_pipe_active_pct"]
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        return {"tflops": float(d["bf16_tflops_sustained"]), "tflops_burst": float(d["bf16_tflops"]),
                "hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    except Exception:
        return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
                "source": "fallback (B200_PROFILING.md: 1.59 PF burst / ~1.4 PF sustained)"}


# ------------------------------------------------------------------------------------------------ CPU arm
def _oracle_modules():
    from b200 import synth
    from oracle import frontend_np, model_torch
    return synth, frontend_np, model_torch


def oracle_pass(waves, vsd, msd):
    """The reference's algorithm on the CPU: numpy float64 front end per clip (vggish_input.py:30-82), torch fp32
    VGGish (vggish.py:21-31), head (model.py:258-269)."""
    _, frontend_np, model_torch = _oracle_modules()
    ex = np.concatenate([frontend_np.waveform_to_examples(w.astype(np.float64)) for w in waves])
    x = torch.from_numpy(ex).float()[:, None]
    with torch.no_grad():
        emb = model_torch.vgg_forward(vsd, x)
        return model_torch.mla_forward(msd, emb.reshape(len(waves), 10, 128), MODEL_CONF)


def time_cpu(vsd, msd, steps, warmup, budget_s):
    """Times `steps` oracle passes (after `warmup`) on a sample sized so that the whole run takes ~budget_s."""
    synth, _, _ = _oracle_modules()
    probe = synth.fast_clips(0, 2).numpy()
    oracle_pass(probe[:1], vsd, msd)                         # page-in / thread pools
    t0 = time.perf_counter()
    oracle_pass(probe, vsd, msd)
    per_clip = (time.perf_counter() - t0) / 2
    n = int(max(1, min(16, budget_s / max(1e-3, per_clip * (steps + warmup)))))
    waves = synth.fast_clips(0, n).numpy()
    for _ in range(warmup):
        oracle_pass(waves, vsd, msd)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_pass(waves, vsd, msd)
    dt = time.perf_counter() - t0
    return {"clips_per_s": n * steps / dt, "ms_per_step": 1e3 * dt / steps, "sample_clips": n, "steps": steps,
            "threads": torch.get_num_threads()}


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    synth, _, _ = _oracle_modules()
    vsd = synth.vggish_state_dict(0)
    msd = synth.mla_state_dict(MODEL_CONF, 128, 600, N_CLASSES, 10, seed=2)
    # ~90 s of CPU work for the whole run by default (VMB_BENCH_CPU_BUDGET_S overrides, e.g. for the CPU test suite)
    r = time_cpu(vsd, msd, args.steps, args.warmup, budget_s=float(os.environ.get("VMB_BENCH_CPU_BUDGET_S", "90")))
    sample = (f"{r['sample_clips']} of the {args.clips} clips of a step per timed pass, {args.steps} passes after "
              f"{args.warmup} warm-ups; numpy float64 front end + torch {torch.__version__} CPU fp32")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["clips_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": r["clips_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["clips_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cpus": os.cpu_count(),
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.stop_flag = threading.Event()
        self.mhz, self.mask, self.max_mhz, self.power, self.ok = [], 0, None, [], False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.mhz.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    m = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    m = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.mask |= int(m)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.ok or not self.mhz:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML sampling unavailable"}
        return {"sm_mhz": float(np.median(self.mhz)), "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if self.mask & b], "samples": len(self.mhz),
                "power_w_max": max(self.power) if self.power else None}


def bind_to_gpu_numa_node(device_index):
    """Pin this rank's threads (and therefore its first-touch pinned host buffers) to the CPUs NVML reports as local
    to its GPU, so that at N > 1 every rank's H2D copies read host memory on the GPU's own socket."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to the GPU"
    except Exception as e:  # noqa: BLE001
        return f"unavailable ({type(e).__name__})"
    return "unavailable"


def workload_config(args, world):
    return {"workload": f"batch {args.clips} synthetic 10 s 16 kHz clips per GPU, waveform -> log-mel -> VGGish -> "
                        f"multi-level attention scores ({N_CLASSES} classes, model_conf [2,1])",
            "clips_per_gpu_per_step": args.clips, "global_clips_per_step": args.clips * world,
            "samples_per_clip": CLIP_SAMPLES, "examples_per_clip": 10, "n_classes": N_CLASSES,
            "parallelism": f"batch-sharded x{world}, no collective on the data path",
            "l2": f"input batch {args.clips * CLIP_SAMPLES * 4 / 1e6:.0f} MB fp32 + >1 GB of activations per step "
                  "(larger than the 126 MB L2)",
            "flop_per_clip": FLOP_PER_CLIP}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, local_rank, world):
    import torch.distributed as dist
    from b200 import _lib, engine, synth

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    engine.require_b200(dev)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    vsd = synth.vggish_state_dict(0)
    msd = synth.mla_state_dict(MODEL_CONF, 128, 600, N_CLASSES, 10, seed=2)
    vgg = engine.VggishHandle(vsd, dev)
    head = engine.MlaHandle(msd, MODEL_CONF, 128, 600, N_CLASSES, 10, dev)
    pipe = engine.Pipeline(vgg, head)

    wave_host = synth.fast_clips(rank * args.clips, args.clips).pin_memory()          # this rank's shard
    scores_host = torch.empty(args.clips, N_CLASSES).pin_memory()
    wave_dev = wave_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    for _ in range(max(3, args.warmup)):
        scores = pipe.forward(wave_dev)
    L.vmb_profile_collect(None, None, 1)
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    L.vmb_profile_enable(1)
    launches0 = L.vmb_launch_count()
    ev0.record()
    for _ in range(args.steps):
        scores = pipe.forward(wave_dev)
    ev1.record()
    barrier()
    launches = L.vmb_launch_count() - launches0
    L.vmb_profile_enable(0)
    sampler.stop_flag.set()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    stage_ms = (torch.zeros(len(STAGES), dtype=torch.float64)).numpy()
    stage_calls = np.zeros(len(STAGES), dtype=np.int64)
    L.vmb_profile_collect(stage_ms.ctypes.data, stage_calls.ctypes.data, 1)
    sampler.join(timeout=1.0)
    finite = bool(torch.isfinite(scores).all().item())

    # ---- end to end through the host-buffer C-ABI entry points (pinned host in, pinned host out).  Every step's
    # H2D copy of its inputs, its compute and its D2H copy of the scores happen inside the timed region; two steps
    # are kept in flight (submit i+1 before collecting i, as a DataLoader-fed loop would), so the copy of step i+1
    # overlaps the compute of step i.  The strictly serial variant (one blocking call per step) is timed as well.
    if args.e2e_microbatch <= 0:
        args.e2e_microbatch = args.clips
    scores_host2 = torch.empty(args.clips, N_CLASSES).pin_memory()
    outs = (scores_host, scores_host2)
    for _ in range(max(3, args.warmup)):
        pipe.forward_host(wave_host, scores_host, clips_per_batch=args.e2e_microbatch)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    pending = pipe.submit_host(wave_host, outs[0], clips_per_batch=args.e2e_microbatch)
    for i in range(1, args.steps):
        nxt = pipe.submit_host(wave_host, outs[i & 1], clips_per_batch=args.e2e_microbatch)
        pipe.wait_host(pending)
        pending = nxt
    pipe.wait_host(pending)
    e1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
    e2e_equal = bool(torch.equal(outs[(args.steps - 1) & 1], scores.cpu()))
    # the same pipelined loop fed with 16-bit PCM (the WAV sample format, vggish_input.py:96-98): half the H2D bytes
    pcm_host = torch.clamp(torch.round(wave_host * 32768.0), -32768, 32767).to(torch.int16).pin_memory()
    pend = pipe.submit_host(pcm_host, outs[0], clips_per_batch=args.e2e_microbatch)
    pipe.wait_host(pend)
    barrier()
    t0 = time.perf_counter()
    pending = pipe.submit_host(pcm_host, outs[0], clips_per_batch=args.e2e_microbatch)
    for i in range(1, args.steps):
        nxt = pipe.submit_host(pcm_host, outs[i & 1], clips_per_batch=args.e2e_microbatch)
        pipe.wait_host(pending)
        pending = nxt
    pipe.wait_host(pending)
    barrier()
    pcm_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe.forward_host(wave_host, scores_host, clips_per_batch=args.serial_microbatch)
    barrier()
    serial_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))

    # ---- latency of BASELINE.json configs[0]: ONE 10 s clip, host buffer in -> scores on the host (blocking call)
    one_host = wave_host[:1].clone().pin_memory()
    one_out = torch.empty(1, N_CLASSES).pin_memory()
    for _ in range(5):
        pipe.forward_host(one_host, one_out, clips_per_batch=1)
    t0 = time.perf_counter()
    for _ in range(50):
        pipe.forward_host(one_host, one_out, clips_per_batch=1)
    b1_ms = 1e3 * (time.perf_counter() - t0) / 50

    # ---- the same device-resident step with the accuracy mode of the VGGish body (hi + lo bf16 planes, 3x the tensor
    # work): the mode whose ranking metric / uint8 output agree with the fp32 reference; reported, not the metric
    acc_steps = max(3, min(10, args.steps))
    vgg_split = engine.VggishHandle(vsd, dev, precision="split")
    pipe_split = engine.Pipeline(vgg_split, head)
    for _ in range(2):
        pipe_split.forward(wave_dev)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(acc_steps):
        pipe_split.forward(wave_dev)
    a1.record()
    barrier()
    acc_ms = max_over_ranks(a0.elapsed_time(a1)) / acc_steps
    vgg_split.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n_ex = args.clips * 10
    ig = slice(2, 10)                                                       # conv2 .. fc3 = the tcgen05 kernel
    ig_ms_per_step = float(stage_ms[ig].sum()) / args.steps
    ig_launches_per_step = int(stage_calls[ig].sum()) // args.steps
    ig_flop_per_step = FLOP_IGEMM_PER_EXAMPLE * n_ex
    peaks = measured_peaks()
    achieved = ig_flop_per_step / (ig_ms_per_step * 1e-3) / 1e12 if ig_ms_per_step > 0 else None
    per_stage = {STAGES[i]: round(float(stage_ms[i]) / args.steps, 4) for i in range(len(STAGES)) if stage_calls[i]}
    stage_flop = dict(zip(STAGES[1:10], [f * n_ex for f in FLOP_CONV + FLOP_FC]))
    stage_tflops = {k: round(stage_flop[k] / (per_stage[k] * 1e-3) / 1e12, 1) for k in stage_flop if per_stage.get(k)}
    clips_per_s = args.clips * world / (ms_step * 1e-3)
    line = {
        "metric": METRIC, "value": clips_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": args.clips * world / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": args.clips * CLIP_SAMPLES * 4, "d2h_bytes_per_step": args.clips * N_CLASSES * 4,
                "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps,
                "microbatch_clips": args.e2e_microbatch, "matches_device_path": e2e_equal, "steps_in_flight": 2,
                "serial_value": args.clips * world / (serial_ms / args.steps * 1e-3),
                "serial_ms_per_step": serial_ms / args.steps, "serial_microbatch_clips": args.serial_microbatch,
                "pcm16_value": args.clips * world / (pcm_ms / args.steps * 1e-3),
                "pcm16_h2d_bytes_per_step": args.clips * CLIP_SAMPLES * 2,
                "api": "vmb_pipeline_submit_host / vmb_pipeline_wait_host via b200.engine.Pipeline (serial_*: one "
                       "blocking vmb_pipeline_forward_host call per step)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": (achieved / peaks["tflops"]) if achieved else None, "traffic": igemm_traffic(),
                     "frac_of_burst_peak": (achieved / peaks["tflops_burst"]) if achieved else None,
                     "tensor_pipe_active_pct_ncu": tensor_pipe_util(),
                     "kernel": "igemm_pair_kernel (tcgen05 cta_group::2 implicit GEMM: conv3_1..conv4_2, fc1, fc2) + "
                               "igemm_bf16_kernel (cta_group::1: conv2, fc3)",
                     "launches_per_step": ig_launches_per_step, "algorithmic_flop_per_launch":
                         ig_flop_per_step / max(1, ig_launches_per_step),
                     "avg_launch_ms": ig_ms_per_step / max(1, ig_launches_per_step), "peak_source": peaks["source"],
                     "whole_step_tflops": FLOP_PER_CLIP * args.clips / (ms_step * 1e-3) / 1e12,
                     "whole_step_frac": FLOP_PER_CLIP * args.clips / (ms_step * 1e-3) / 1e12 / peaks["tflops"]},
        "stage_ms_per_step": per_stage, "stage_tflops": stage_tflops,
        "single_clip_latency_ms": b1_ms, "host_affinity": numa,
        "accuracy_mode": {"value": args.clips * world / (acc_ms * 1e-3), "unit": UNIT, "ms_per_step": acc_ms,
                          "dtype": "split bf16 (hi + lo planes, fp32-class body)"},
        "scores_finite": finite,
    }
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu(vsd, msd, steps=3, warmup=1, budget_s=20.0)
        line["cpu_baseline"] = {
            "value": r["clips_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
            "sample": f"{r['sample_clips']} of the {args.clips} clips per pass, 3 timed passes after 1 warm-up "
                      f"(numpy float64 front end + torch CPU fp32 VGGish + head); host has {os.cpu_count()} CPUs"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    global _OUT
    _OUT = _json_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--clips", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--e2e-microbatch", type=int, default=0,
                    help="clips per H2D/compute micro-batch in the pipelined e2e leg (0 = the whole step)")
    ap.add_argument("--serial-microbatch", type=int, default=64,
                    help="clips per micro-batch in the serial (one blocking call per step) e2e variant")
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
