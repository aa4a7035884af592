#!/usr/bin/env python
"""bench.py -- VQ-encoded patches/sec of the fused vector-quantisation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the standalone VectorQuantizer forward on 2^24
synthetic patch vectors per GPU, repo-default codebook K=256, D=32 (z = 0.1*randn, seed
1234+rank; codebook U(-1/K, 1/K), seed 0).  One step = one pass of the hot path over that
batch: distance + argmin + gather + straight-through value + loss + histogram/perplexity
(z_q (N,D) fp32 and int64 ids written; the (N,K) one-hot is not materialised -- SURVEY.md
section 8(d)).  Weak scaling: every rank quantises its own shard, the codebook is replicated,
the only collective is the K-element all-reduce of the code histogram.

One JSON line on stdout (rank 0).  `value` = device-timed whole-job patches/s with inputs
resident in HBM; `e2e` = the same through the host-buffer C-ABI call (pinned host arrays in,
z_q + ids + scalars back on the host, copies inside the timed region); `roofline` = the
dominant kernel against the measured HBM peak; `cpu_baseline` = the reference's op sequence
on this box's host cores (oracle port, bounded sample).

Further objects on the same line (each measured in this run, none part of `value`): `bulk_encode` = BASELINE
configs[2], bulk latent-dataset encoding through LatentSpaceEncoder (synthetic cycles in, ids out, fused encoder chain
+ fused quantiser) with its tensor-roofline fraction, id match against the fp32 encoder and the reference's loop on
host cores, and `bulk_encode.e2e`: the same workload through create_latent_space_dataset_VQ_VAE_IDs from loader batches in
host memory to the ids on the host (+ at N = 1 the reference's overlapping-window geometry, de-duplicated and not); `config1_forward` = BASELINE configs[0], the whole VQVAEPatch.forward at batch 256 on the GPU next to the
reference's op sequence on host cores; `backward` = the straight-through backward call.

--impl reference times that CPU port alone (the reference is a Python/PyTorch program whose
own files cannot travel to the GPU box) on the same config and metric.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_CODES, DIM, BETA = 256, 32, 0.25
N_VECTORS = 1 << 24
METRIC = "vq_encoded_patches_per_sec"
UNIT = "patches/s"
FALLBACK_HBM_GBS = 6650.0
_emit = None      # set by main(): writes the JSON line to the real stdout


def workload_config(n_gpus: int, n_rows: int) -> dict:
    return {
        "workload": f"VectorQuantizer.forward, K={K_CODES} D={DIM} (repo default), N={n_rows} vectors per GPU "
                    f"(BASELINE configs[1])",
        "codebook": "U(-1/K,1/K) seed 0", "inputs": "0.1*randn seed 1234+rank",
        "outputs": "z_q fp32 (N,D) + int64 ids + loss + perplexity + histogram; no (N,K) one-hot",
        "parallelism": f"batch-sharded x{n_gpus}, codebook replicated",
        "l2": f"inputs larger than L2 ({n_rows * DIM * 4 / 2**20:.0f} MiB read per step)",
    }


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed regions
# ---------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, period: float = 0.02):
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thread = None
        self._index, self._period = index, period
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _reason_names(self, mask: int):
        nv = self._nv
        table = [("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"),
                 ("hw_power_brake", "nvmlClocksThrottleReasonHwPowerBrakeSlowdown"),
                 ("sync_boost", "nvmlClocksThrottleReasonSyncBoost"),
                 ("app_clocks", "nvmlClocksThrottleReasonApplicationsClocksSetting")]
        return [name for name, attr in table if mask & int(getattr(nv, attr, 0))]

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                getter = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.reasons.update(self._reason_names(int(getter(self._h))))
            except Exception:
                pass
            self._stop.wait(self._period)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                # NVML's board power is a ~1 s moving average: over a 22 ms device loop plus the PCIe-bound e2e leg it
                # stays far below what the kernels draw back to back (tools/clock_probe.py: 990 W, sw_power_cap)
                "power_w_nvml_avg": statistics.median(self.power) if self.power else None,
                "power_w_nvml_avg_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------
# the reference's CPU path (oracle port): cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------
def cpu_port_rate(rows_per_call: int, calls: int, warm: int):
    """Times the reference's op sequence (oracle.torch_port_forward; model/vector_quantizer.py:
    88-119) on host cores.  Returns (patches/s, seconds per call list, threads)."""
    import torch
    from oracle import vq_oracle as O

    torch.set_float32_matmul_precision("highest")
    try:   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core it may run on
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    g = torch.Generator().manual_seed(1234)
    z = 0.1 * torch.randn(rows_per_call, DIM, generator=g)
    g0 = torch.Generator().manual_seed(0)
    weight = (torch.rand(K_CODES, DIM, generator=g0) * 2 - 1) / K_CODES
    times = []
    with torch.no_grad():
        for i in range(warm + calls):
            t0 = time.perf_counter()
            O.torch_port_forward(z, weight, BETA)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return rows_per_call * len(times) / sum(times), times, torch.get_num_threads()


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: size one step so that warmup+steps finish within ~2 minutes
    probe_rows = 1 << 16
    rate, _, threads = cpu_port_rate(probe_rows, 2, 1)
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    rows = int(min(1 << 20, max(1 << 14, rate * budget_s)))
    rows = 1 << (rows.bit_length() - 1)
    rate, times, threads = cpu_port_rate(rows, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    sample = f"{rows} of the {N_VECTORS} vectors per step (2^{rows.bit_length() - 1}-vector chunk), scaled per vector"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, N_VECTORS),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference's op sequence (model/vector_quantizer.py:88-119) as torch CPU ops on all host threads; "
                "the reference's own files are not present on the GPU box (oracle port, DESIGN.md)",
    }
    _emit(line)


# ---------------------------------------------------------------------------------------
# host-side helpers of the e2e leg
# ---------------------------------------------------------------------------------------
def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so that the pinned staging buffers it
    allocates next are first-touched on that GPU's NUMA node (eight ranks otherwise share one node's memory
    controllers).  Returns (previous affinity, number of CPUs now used) or (None, None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i * 64 + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        before = os.sched_getaffinity(0)
        cpus &= before
        if cpus:
            os.sched_setaffinity(0, cpus)
            return before, len(cpus)
    except Exception:
        pass
    return None, None


def copy_roofline(torch, dev, z_host, zq_host, idx_host, reps: int = 3):
    """Bare pinned copies of the e2e leg's buffers, no kernels: H2D alone, D2H alone, and both at once on two streams
    (what vqb_encode_host overlaps).  The duplex time is the floor of the e2e step on this host."""
    d_in = torch.empty(z_host.shape, dtype=z_host.dtype, device=dev)
    d_zq = torch.empty(zq_host.shape, dtype=zq_host.dtype, device=dev)
    d_idx = torch.empty(idx_host.shape, dtype=idx_host.dtype, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(h2d: bool, d2h: bool) -> float:
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream(dev))
        s2.wait_stream(torch.cuda.current_stream(dev))
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(z_host, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    zq_host.copy_(d_zq, non_blocking=True)
                    idx_host.copy_(d_idx, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(s1)
        torch.cuda.current_stream(dev).wait_stream(s2)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    timed(True, True)
    h2d_ms, d2h_ms, duplex_ms = timed(True, False), timed(False, True), timed(True, True)
    in_b = z_host.numel() * 4
    out_b = zq_host.numel() * 4 + idx_host.numel() * idx_host.element_size()
    return {"h2d_gbs": in_b / h2d_ms / 1e6, "d2h_gbs": out_b / d2h_ms / 1e6, "duplex_ms": duplex_ms,
            "h2d_ms": h2d_ms, "d2h_ms": d2h_ms}


# ---------------------------------------------------------------------------------------
# BASELINE configs[2]: bulk latent-dataset encoding (dataloader/latentspace_dataloader.py:205-243)
# ---------------------------------------------------------------------------------------
CYCLE_FLOP = 16 * 8_463_360        # patchify 25.6 k + res blocks 8388.6 k + proj 32.8 k + VQ 16.4 k per token


def default_model(torch, vqb200, dev):
    torch.manual_seed(0)
    return vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=K_CODES, embedding_dim=DIM, n_resblocks=8,
                             learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()


def overlapping_windows_leg(torch, enc, model, n_cycles: int = 50_000, seq_len: int = 20, batch: int = 512):
    """The reference's data-set geometry: windows of 20 cycles with a stride of ONE cycle over a stream of cycles
    (dataloader/asimow_dataloader.py:185-206), loader batches of 512 windows (pageable: strided views of the host stream).
    Patches are counted the way the reference produces them (every cycle of every window).  Single-GPU runs only (the
    staging copies of pageable batches use every host thread)."""
    import numpy as np
    T = int(model.enc_out_len)
    stream = torch.randn(n_cycles * 200, 2, generator=torch.Generator().manual_seed(7))
    n_windows = n_cycles - seq_len + 1
    windows = stream.as_strided((n_windows, seq_len * 200, 2), (200 * 2, 2, 1))
    loader = [windows[i:i + batch] for i in range(0, n_windows, batch)]
    res, ref = {"cycles": n_cycles, "windows": n_windows, "loader_batches": len(loader), "unit": UNIT}, None

    def best_of(fn, reps=3):
        best, val = None, None
        for rep in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            val = fn()
            dt = time.perf_counter() - t0
            best = dt if best is None or (rep and dt < best) else best
        return best, val

    keep = enc.dedupe
    for name, mode in (("every_window_encoded", False), ("dedupe_per_call", True), ("dedupe_per_data_set", "dataset")):
        enc.dedupe = mode
        sec, (ids, _) = best_of(lambda: enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True,
                                                                                  no_labels=True))
        ref = ids if ref is None else ref
        res[name] = {"value": n_windows * seq_len * T / sec, "ms": sec * 1e3, "ids_equal": bool(np.array_equal(ids, ref))}
    enc.dedupe = keep
    sec, (ids, _) = best_of(lambda: enc.create_latent_space_dataset_from_cycles(stream.view(n_cycles, 200, 2), None, seq_len=seq_len,
                                                                                has_patch_embed=True))
    res["from_cycle_stream"] = {"value": ids.shape[0] * seq_len * T / sec, "ms": sec * 1e3, "windows": int(ids.shape[0]),
                                "ids_equal": bool(np.array_equal(ids, ref[: ids.shape[0]]))}
    res["what"] = ("create_latent_space_dataset_VQ_VAE_IDs over the windows with dedupe off / True / 'dataset' (vqb_row_keys + "
                   "vqb_dedupe_first: every distinct cycle of a call / of the data set encoded once), and "
                   "create_latent_space_dataset_from_cycles on the stream itself (every cycle encoded once, windows as a sliding "
                   "view of the ids; the reference's create_sequence_ds keeps n - seq_len windows); wall clock, best of 2 after a warm-up")
    return res


def bulk_e2e_leg(torch, dist, enc, model, dev, rank, world, barrier, resident_ms_per_cycle):
    """configs[2] end to end through the reference-facing call: loader batches of windows in PINNED HOST memory ->
    LatentSpaceEncoder.create_latent_space_dataset_VQ_VAE_IDs -> the int64 ids as ONE numpy array on the host.  The timed
    region (wall clock, device idle on both sides) holds every host->device copy of the cycles, the encoder and quantiser
    launches, the device->host copies of the ids and the host-side assembly of the array.  Two batch geometries: the
    reference's own (512 windows of 20 cycles per loader batch, train_transformer_mtasks.py:214,
    dataloader/latentspace_dataloader.py:225-238) and large batches (4096 windows of 16 cycles = the resident leg's chunk).
    All ranks run at once on their own data; the time is the maximum over ranks."""
    T = int(model.enc_out_len)
    out = {}
    g = torch.Generator(device=dev).manual_seed(2000 + rank)
    for name, windows, seq_len, n_batches, pinned in (("reference_batches", 512, 20, 52, True),
                                                       ("large_batches", 4096, 16, 8, True),
                                                       ("large_batches_pageable", 4096, 16, 8, False)):
        if not pinned and world > 1:
            continue                          # the staging copy of pageable batches uses every host thread: one rank only
        hosts = []
        for _ in range(3):                    # three distinct host batches, cycled through the loader list
            h = torch.empty(windows, seq_len * 200, 2, dtype=torch.float32, pin_memory=pinned)
            h.copy_(torch.randn(windows, seq_len * 200, 2, device=dev, generator=g))
            hosts.append(h)
        loader = [hosts[i % 3] for i in range(n_batches)]
        times = []
        for rep in range(3):                  # first pass untimed (pinned staging buffers of the id writer, allocator)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            ids, _ = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True, no_labels=True)
            dt = time.perf_counter() - t0
            if rep:
                times.append(dt)
        assert ids.shape == (windows * n_batches, seq_len, T) and str(ids.dtype) == "int64"
        t = torch.tensor([min(times)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
        cycles = windows * seq_len * n_batches
        out[name] = {"value": world * cycles * T / sec, "unit": UNIT, "per_gpu": cycles * T / sec,
                     "ms_per_data_set": sec * 1e3, "cycles_per_gpu": cycles, "loader_batch": [windows, seq_len * 200, 2],
                     "loader_batches": n_batches, "host_memory": "pinned" if pinned else "pageable (staged through pinned buffers)",
                     "h2d_bytes_per_data_set": cycles * 200 * 2 * 4, "d2h_bytes_per_data_set": cycles * T * 8,
                     "frac_of_resident_rate": resident_ms_per_cycle * 1e-3 * cycles / sec}
        del hosts, loader, ids
    if world == 1:        # (pageable loader batches: their staging copies use every host thread, torchrun gives a rank one)
        out["overlapping_windows"] = overlapping_windows_leg(torch, enc, model)
    out["what"] = ("wall clock around LatentSpaceEncoder.create_latent_space_dataset_VQ_VAE_IDs(loader, no_labels=True): host batches "
                   "in, one int64 numpy array out; H2D of batch i + 1 on a copy stream under the encode of batch i, ids back "
                   "through pinned buffers on a second copy stream; best of 2 timed passes after one warm-up, max over ranks")
    return out


def bulk_encode_leg(args, torch, dist, vqb200, lib, dev, rank, world, barrier):
    from vqb200.dataloader import LatentSpaceEncoder
    chunk, n_chunks = args.bulk_chunk, args.bulk_chunks
    model = default_model(torch, vqb200, dev)
    model.encoder_mode = "fused_bf16"
    enc = LatentSpaceEncoder(model, window_size=200, device=str(dev))
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    pool = [torch.randn(chunk, 200, 2, device=dev, generator=g) for _ in range(2)]     # 2 x 105 MB: larger than L2
    with torch.no_grad():
        for i in range(2):
            ids = enc.get_latent_space_IDs(pool[i & 1], has_patch_embed=True)
        barrier()
        l0 = lib.vqb_launch_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_chunks):
            ids = enc.get_latent_space_IDs(pool[i & 1], has_patch_embed=True)
        e1.record()
        barrier()
        launches = lib.vqb_launch_counter() - l0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        # id match of the bf16-operand encoder against the fp32 encoder (TF32 off) on a sample of the same cycles
        sample = pool[0][:4096]
        got = enc.get_latent_space_IDs(sample, has_patch_embed=True).view(-1)
        tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        model.encoder_mode = "torch"
        ref = enc.get_latent_space_IDs(sample, has_patch_embed=True).view(-1)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        match = float((got == ref).float().mean().item())
        # the fp32-faithful fused encoder (VQVAEPatch's default for inference): one chunk timed, ids against the same fp32 ids
        model.encoder_mode = "fused_fp32"
        got32 = enc.get_latent_space_IDs(sample, has_patch_embed=True).view(-1)
        enc.get_latent_space_IDs(pool[1], has_patch_embed=True)
        torch.cuda.synchronize()
        l1 = lib.vqb_launch_counter()
        e0.record()
        enc.get_latent_space_IDs(pool[0], has_patch_embed=True)
        e1.record()
        torch.cuda.synchronize()
        fp32_ms, fp32_launches = e0.elapsed_time(e1), lib.vqb_launch_counter() - l1
        match32 = float((got32 == ref).float().mean().item())
        model.encoder_mode = "fused_bf16"
        e2e = bulk_e2e_leg(torch, dist, enc, model, dev, rank, world, barrier, ms_max / n_chunks / chunk)
    if rank != 0:
        return None
    cycles = world * chunk * n_chunks
    patches = cycles * model.enc_out_len
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak_tf, peak_src = float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
    achieved_tf = cycles * CYCLE_FLOP / (ms_max * 1e-3) / 1e12 / world       # per GPU
    chain_traffic = None
    try:        # DRAM bytes of the dominant kernel (the fused encoder launch) per launch, from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            chain_traffic = float(json.load(f)["enc_chain_bytes_per_token"]) * chunk * model.enc_out_len
    except Exception:
        pass
    out = {
        "workload": f"BASELINE configs[2]: LatentSpaceEncoder.get_latent_space_IDs on synthetic cycles randn(., 200, 2), "
                    f"{n_chunks} chunks of {chunk} cycles per GPU resident in HBM, ids (int64) out; default model "
                    f"(H=512, 8 res blocks, K={K_CODES}, D={DIM}), random-init weights seed 0",
        "value": patches / (ms_max * 1e-3), "unit": UNIT, "per_gpu": patches / world / (ms_max * 1e-3),
        "n_gpus": world, "ms_per_chunk": ms_max / n_chunks, "cycles_per_s": cycles / (ms_max * 1e-3),
        "encoder_precision": "encoder_mode='fused_bf16': vqb_patch_split + vqb_encoder_chain (one launch: patch embedding "
                             "and projection with bf16 hi+lo operand pairs = fp32-accurate to 2^-16, the 16 hidden layers "
                             "with bf16 operands, fp32 accumulation and fp32 residual stream), exact quantiser",
        "id_match_vs_fp32_encoder": {"rate": match, "rows": int(got.numel())},
        "fp32_faithful": {"mode": "encoder_mode='fused_fp32' (VQVAEPatch's inference default): vqb_patch_embed (fp32 FMA) + "
                                  "vqb_token_pair + 17 x vqb_token_linear_split (bf16 hi+lo operand pairs, three tcgen05 products "
                                  "per layer, fp32 accumulation, erf GELU, fp32 residual stream), exact quantiser; rank 0, one chunk",
                          "per_gpu": chunk * model.enc_out_len / (fp32_ms * 1e-3), "unit": UNIT, "ms_per_chunk": fp32_ms,
                          "gpu_launches_per_chunk": fp32_launches,
                          "id_match_vs_fp32_encoder": {"rate": match32, "rows": int(got32.numel())},
                          "issued_tflops": 3.0 * chunk * CYCLE_FLOP / (fp32_ms * 1e-3) / 1e12},
        "gpu_launches_per_chunk": launches / n_chunks,
        "e2e": e2e,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "flop_per_cycle": CYCLE_FLOP, "peak_source": peak_src,
                     "traffic": chain_traffic, "kernel": "enc_chain_kernel<512> (vqb_encoder_chain: one launch per chunk)"},
    }
    if not args.no_cpu:
        # the reference's loop on host cores: per 512-window batch, cycle by cycle (one encode call per cycle slice)
        from oracle import vq_oracle as O
        torch.set_float32_matmul_precision("highest")
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))
        except Exception:
            pass
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        xb = torch.randn(512, 200, 2, generator=torch.Generator().manual_seed(1000))
        times = []
        with torch.no_grad():
            for i in range(4):
                t0 = time.perf_counter()
                z_e = O.torch_port_encode(sd, xb.clone(), 25)
                O.torch_port_forward(z_e, sd["vector_quantization.embedding.weight"], BETA)[4].numpy().reshape(512, -1)
                dt = time.perf_counter() - t0
                if i:
                    times.append(dt)
                if sum(times) > 15.0:
                    break
        out["cpu_baseline"] = {"value": 512 * 16 * len(times) / sum(times), "unit": UNIT, "cores": torch.get_num_threads(),
                               "kind": "port",
                               "sample": f"{len(times)} encode calls of 512 cycles (the reference's per-cycle-slice call, "
                                         f"dataloader/latentspace_dataloader.py:231-235) in {sum(times):.1f} s"}
    return out


def sweep_corners_leg(args, torch, vqb200, dev, n):
    """A few corners of BASELINE configs[1]'s K x D sweep (tools/sweep.py has the whole grid): the large-codebook and
    wide-vector shapes run on the tile-stationary kernel (csrc/vq_fwd_tcs.cu)."""
    from vqb200 import ops
    rows = []
    for K, D in ((512, 32), (1024, 32), (4096, 32), (256, 64), (256, 128)):
        z = 0.1 * torch.randn(n, D, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
        w = ((torch.rand(K, D, generator=torch.Generator().manual_seed(0)) * 2 - 1) / K).to(dev)
        res = {"K": K, "D": D}
        for name, kw in (("full_ms", {}), ("ids_only_ms", {"want_zq": False, "want_loss": False})):
            for _ in range(2):
                ops.forward(z, w, BETA, **kw)
            reps = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ops.forward(z, w, BETA, **kw)
            e1.record()
            torch.cuda.synchronize(dev)
            res[name] = e0.elapsed_time(e1) / reps
        res["hbm_gbs_full"] = n * (8 * D + 8) / res["full_ms"] / 1e6
        res["algorithmic_tflops_ids_only"] = 2.0 * K * D * n / res["ids_only_ms"] / 1e9
        rows.append(res)
        del z
    return {"N": n, "what": "ms per vqb_forward call (all outputs / ids only), device-timed, 3 calls after 2 warm-ups", "rows": rows}


def config1_leg(args, torch, vqb200, dev):
    """BASELINE configs[0]: VQVAEPatch reconstruction forward, batch 256, fp32, eval."""
    model = default_model(torch, vqb200, dev)
    x = torch.randn(256, 200, 2, generator=torch.Generator().manual_seed(0))
    xd = x.to(dev)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_forward():
        with torch.no_grad():
            for _ in range(2):
                res = model(xd)
            e0.record()
            for _ in range(5):
                model(xd)
            e1.record()
            torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / 5, res

    model.encoder_mode = model.decoder_mode = "torch"          # the stock PyTorch layers around the fused quantiser
    torch_ms, (_, hat_torch, _) = timed_forward()
    model.encoder_mode = model.decoder_mode = "auto"           # the module's default: fp32-faithful tcgen05 layers for inference
    gpu_ms, (_, hat_auto, _) = timed_forward()
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    out = {"workload": "BASELINE configs[0]: VQVAEPatch.forward (encode + quantise + decode), batch 256, fp32 (TF32 off), eval",
           "gpu_ms": gpu_ms, "gpu_patches_per_s": 256 * 16 / (gpu_ms * 1e-3),
           "modes": "encoder_mode = decoder_mode = 'auto' (the default) = 'fused_fp32': vqb_token_linear_split / vqb_token_conv_split "
                    "(bf16 hi+lo operand pairs, three tcgen05 products per layer, fp32 accumulation and residual streams, erf GELU), "
                    "exact quantiser",
           "x_hat_max_abs_dev_vs_torch_layers": float((hat_auto - hat_torch).abs().max().item()),
           "x_hat_max_abs": float(hat_torch.abs().max().item()),
           "torch_layers": {"gpu_ms": torch_ms, "gpu_patches_per_s": 256 * 16 / (torch_ms * 1e-3),
                            "modes": "encoder_mode = decoder_mode = 'torch' (stock PyTorch fp32 layers, fused quantiser)"}}
    # the same forward with BOTH halves on the hand-written tcgen05 layer kernels (bf16 operands, fp32 accumulation and
    # residual streams; the quantiser stays exact), at the configuration's batch and at a batch that fills the GPU
    with torch.no_grad():
        model.encoder_mode = model.decoder_mode = "fused_bf16"
        fused = {}
        for name, xin in (("batch_256", xd), ("batch_8192", torch.randn(8192, 200, 2, device=dev,
                                                                        generator=torch.Generator(device=dev).manual_seed(1)))):
            for _ in range(2):
                _, hat, _ = model(xin)
            e0.record()
            for _ in range(5):
                model(xin)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / 5
            fused[name] = {"gpu_ms": ms, "gpu_patches_per_s": xin.shape[0] * 16 / (ms * 1e-3)}
        # decoder precision on IDENTICAL quantised latents (the two encoders differ on ~0.14 % of the ids, which would
        # dominate a comparison of whole forwards): fused decoder against the fp32 PyTorch modules
        z_q = model.vector_quantization(model.encode(xd))[1]
        got = model.decode(z_q)
        model.decoder_mode = "torch"
        want = model.decode(z_q)
        model.decoder_mode = "fused_bf16"
        fused["decoder_max_abs_dev_vs_fp32"] = float((got - want).abs().max().item())
        fused["decoder_out_max_abs"] = float(want.abs().max().item())
        fused["precision"] = "encoder_mode = decoder_mode = 'fused_bf16' (vqb_encoder_chain, vqb_token_conv, vqb_token_out_proj)"
        fused["decoder_flop_per_token"] = 2 * 32 * 512 + 16 * 2 * 1536 * 512 + 2 * 512 * 2560 + 5 * 2 * 512 * 5
        model.encoder_mode = model.decoder_mode = "torch"
    out["fused_bf16"] = fused
    if not args.no_cpu:
        from oracle import vq_oracle as O
        torch.set_float32_matmul_precision("highest")
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))
        except Exception:
            pass
        cpu_model = default_model(torch, vqb200, torch.device("cpu"))     # stock PyTorch modules: the decoder half
        sd = {k: v.detach() for k, v in cpu_model.state_dict().items()}
        times = []
        with torch.no_grad():
            for i in range(4):
                t0 = time.perf_counter()
                z_e = O.torch_port_encode(sd, x, 25)                                         # :158-160
                _, z_q, _, _, _ = O.torch_port_forward(z_e, sd["vector_quantization.embedding.weight"], BETA)   # :161
                cpu_model.reverse_patch_embed(cpu_model.decoder(z_q.permute(0, 2, 1)))       # :164-165
                dt = time.perf_counter() - t0
                if i:
                    times.append(dt)
        cpu_ms = 1e3 * sum(times) / len(times)
        out.update({"cpu_ms": cpu_ms, "cpu_patches_per_s": 256 * 16 / (cpu_ms * 1e-3), "cpu_cores": torch.get_num_threads(),
                    "cpu_kind": "port (reference op sequence: encoder loop + VQ ops + stock decoder modules)"})
    return out


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import vqb200
    from vqb200 import ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = vqb200._lib.load()
    n = args.rows

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    z = 0.1 * torch.randn(n, DIM, device=dev, generator=g)
    g0 = torch.Generator().manual_seed(0)
    weight = ((torch.rand(K_CODES, DIM, generator=g0) * 2 - 1) / K_CODES).to(dev)
    if world > 1:
        dist.broadcast(weight, 0)      # replicated codebook

    pending = []

    def step():
        out = ops.forward(z, weight, BETA, path=args.path)
        if world > 1:
            # global code histogram: the path's only collective.  Issued asynchronously (NCCL's own stream) and
            # collected with finish() inside the timed region: the 2 KB all-reduce overlaps the next step's kernels
            # instead of making every rank wait for the slowest one once per step.
            pending.append((dist.all_reduce(out[4], async_op=True), out[4]))
        return out

    def finish():
        for work, _ in pending:
            work.wait()
        pending.clear()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        out = step()
    finish()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.vqb_profile_enable(1)
    launches0 = lib.vqb_launch_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    finish()
    ev1.record()
    barrier()
    launches = lib.vqb_launch_counter() - launches0
    lib.vqb_profile_enable(0)
    ms_total = ev0.elapsed_time(ev1)
    kms, kn = ctypes.c_double(), ctypes.c_int()
    vqb200._lib.check(lib.vqb_profile_collect(ctypes.byref(kms), ctypes.byref(kn)), "vqb_profile_collect")
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    loss, zq, ppl, idx, counts = out

    # ---- end to end through the host-buffer C-ABI call --------------------------------
    e2e = None
    if not args.no_e2e:
        affinity_before, numa_cpus = bind_to_gpu_numa(local_rank)
        z_host = torch.empty((n, DIM), dtype=torch.float32).pin_memory()
        z_host.copy_(z)
        zq_host = torch.empty((n, DIM), dtype=torch.float32).pin_memory()
        idx_host = torch.empty((n,), dtype=torch.int64).pin_memory()
        idx8_host = torch.empty((n,), dtype=torch.uint8).pin_memory()
        counts_host = np.zeros(K_CODES, np.uint64)
        roof = copy_roofline(torch, dev, z_host, zq_host, idx_host)
        enc = ops.HostEncoder(weight, device=local_rank, chunk_rows=args.chunk_rows, depth=3)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            enc.encode(z_host, BETA, zq_out=zq_host, idx_out=idx_host, counts_out=counts_host, path=args.path)
        barrier()
        dev_ms, wall = 0.0, 0.0
        ms_buf = ctypes.c_float()
        for _ in range(e2e_steps):
            t0 = time.perf_counter()
            h_loss, h_ppl = enc.encode(z_host, BETA, zq_out=zq_host, idx_out=idx_host, counts_out=counts_host,
                                       path=args.path)
            wall += time.perf_counter() - t0
            lib.vqb_host_last_ms(enc._ctx, ctypes.byref(ms_buf))
            dev_ms += ms_buf.value
        barrier()
        e2e_launches = enc.last_launches
        # ids-only variant: what dataloader/latentspace_dataloader.py:160-161 actually consumes
        ids_ms = 0.0
        for _ in range(e2e_steps):
            enc.encode(z_host, BETA, idx_out=idx_host, path=args.path)
            lib.vqb_host_last_ms(enc._ctx, ctypes.byref(ms_buf))
            ids_ms += ms_buf.value
        barrier()
        # compact ids: the same ids as one byte each (K <= 256), what a tokeniser that keeps uint8 tokens reads back
        ids8_ms = 0.0
        for _ in range(e2e_steps):
            enc.encode(z_host, BETA, idx_out=idx8_host, path=args.path)
            lib.vqb_host_last_ms(enc._ctx, ctypes.byref(ms_buf))
            ids8_ms += ms_buf.value
        barrier()
        te = torch.tensor([dev_ms, wall * 1e3, ids_ms, ids8_ms, roof["duplex_ms"], roof["h2d_ms"], roof["d2h_ms"]],
                          device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, ids_ms, ids8_ms, duplex_ms, h2d_ms, d2h_ms = (float(v) for v in te.tolist())
        assert torch.equal(idx_host.to(dev), idx.view(-1)), "host path and device path disagree"
        assert torch.equal(idx8_host.to(dev).to(torch.int64), idx.view(-1)), "compact ids disagree"
        e2e = {
            "value": world * n * e2e_steps / (max(dev_ms, 1e-9) * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": n * DIM * 4,
            "d2h_bytes_per_step": n * DIM * 4 + n * 8 + 8 + K_CODES * 8,
            "ms_per_step": dev_ms / e2e_steps, "wall_ms_per_step": wall_ms / e2e_steps, "steps": e2e_steps,
            "timing": "CUDA events inside vqb_encode_host, first H2D to last D2H; wall clock alongside",
            "chunk_rows": args.chunk_rows, "gpu_launches_per_step": e2e_launches,
            "ids_only": {"value": world * n * e2e_steps / (max(ids_ms, 1e-9) * 1e-3), "unit": UNIT,
                         "d2h_bytes_per_step": n * 8 + 8,
                         "copy_floor_frac": h2d_ms / (ids_ms / e2e_steps) if ids_ms > 0 else None},
            "ids_u8": {"value": world * n * e2e_steps / (max(ids8_ms, 1e-9) * 1e-3), "unit": UNIT,
                       "d2h_bytes_per_step": n + 8,
                       "copy_floor_frac": h2d_ms / (ids8_ms / e2e_steps) if ids8_ms > 0 else None},
            # bare pinned copies of the same buffers on this host with all ranks copying at once (max over ranks): the
            # floor of the step; frac = floor / measured step
            "copy_roofline": {"h2d_ms": h2d_ms, "d2h_ms": d2h_ms, "duplex_ms": duplex_ms,
                              "h2d_gbs_per_gpu": n * DIM * 4 / h2d_ms / 1e6,
                              "d2h_gbs_per_gpu": (n * DIM * 4 + n * 8) / d2h_ms / 1e6,
                              "frac": duplex_ms / (dev_ms / e2e_steps) if dev_ms > 0 else None,
                              "what": "concurrent cudaMemcpyAsync H2D (z) + D2H (z_q, ids) of the step's pinned buffers, "
                                      "no kernels, all ranks at once"},
            "numa": {"bound_to_gpu_local_cpus": numa_cpus},
        }
        enc.close()
        del z_host, zq_host, idx_host, idx8_host
        if affinity_before is not None:
            os.sched_setaffinity(0, affinity_before)
    clocks = sampler.stop()

    # ---- straight-through backward on the same batch (reported beside the headline, not part of it) ----
    bwd = None
    if not args.no_bwd:
        gq = torch.randn(n, DIM, device=dev, generator=g)
        gl = torch.tensor(1.0, device=dev)
        for _ in range(3):
            ops.backward(gq, gl, z, idx, weight, BETA)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 10))
        b0.record()
        for _ in range(reps):
            ops.backward(gq, gl, z, idx, weight, BETA)
        b1.record()
        barrier()
        tb = torch.tensor([b0.elapsed_time(b1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        bwd_ms = float(tb.item())
        del gq

    # ---- BASELINE configs[2]: bulk latent-dataset encoding, every rank on its own shard ----
    bulk = None
    if not args.no_bulk:
        idx_keep = idx[: 1 << 16].clone()
        z_keep = z[: 1 << 16].clone()
        del z, zq, idx, out            # make room: the encoder's activations want ~10 GB per chunk
        torch.cuda.empty_cache()
        bulk = bulk_encode_leg(args, torch, dist, vqb200, lib, dev, rank, world, barrier)
        z, idx = z_keep, idx_keep

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cfg1 = config1_leg(args, torch, vqb200, dev) if not args.no_config1 else None
    corners = sweep_corners_leg(args, torch, vqb200, dev, n) if (rank == 0 and not args.no_sweep) else None

    # ---- index match rate on a sample (outside every timed region): against the C oracle and against the
    # reference's own torch op sequence (CPU), with the mismatches a near-tie explains counted separately ------
    from oracle import vq_oracle as O
    sample_rows = 1 << 16
    zs = z[:sample_rows].cpu()
    ours = idx[:sample_rows].view(-1).cpu().numpy()
    ora = O.forward(zs.numpy(), weight.cpu().numpy(), BETA)
    match = float((ours == ora.indices.reshape(-1)).mean())
    torch.set_float32_matmul_precision("highest")
    with torch.no_grad():
        port_idx = O.torch_port_forward(zs, weight.cpu(), BETA)[4].view(-1).numpy()
    port_match = float((ours == port_idx).mean())
    expl = O.explain_mismatches(zs.numpy(), weight.cpu().numpy(), ours, port_idx)

    peak, peak_src = measured_peaks()
    algo_bytes = n * (8 * DIM + 8)                     # read z, write z_q, int64 idx
    k_ms = kms.value / max(1, kn.value)
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get(f"K{K_CODES}_D{DIM}_N{n}_{args.path}")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "kernel": "fused VQ forward (distance+argmin+gather+loss+histogram)", "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms * kn.value / ms_total if ms_total > 0 else None,
                "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src}

    if not args.no_bwd:
        bwd_bytes = n * (12 * DIM + 8)                 # read g_zq, z, idx; write grad_z
        bwd = {"ms_per_call": bwd_ms, "patches_per_s": world * n / (bwd_ms * 1e-3),
               "roofline": {"bound": "hbm", "achieved": bwd_bytes / (bwd_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": bwd_bytes / (bwd_ms * 1e-3) / 1e9 / peak if peak else None,
                            "algorithmic_bytes_per_call": bwd_bytes},
               "what": "vqb_backward (grad_z + dense grad_E), whole call timed with CUDA events: memset + kernel + scale kernel"}

    cpu = None
    if not args.no_cpu:
        # bounded sample of the same workload: 2^18-vector chunks for about 10-20 s of CPU work
        rows = 1 << 18
        rate, times, threads = cpu_port_rate(rows, 1, 1)
        calls = int(max(2, min(64, 12.0 / max(times[0], 1e-3))))
        rate, times, threads = cpu_port_rate(rows, calls, 0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{calls} x {rows}-vector chunks of the {n}-vector workload ({sum(times):.1f} s)"}

    line = {
        "metric": METRIC, "value": world * n * args.steps / (ms_total * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, n), "path": args.path,
        "filter": ("tf32 single product (default for 16 < D <= 32)" if os.environ.get("VQB_TF32", "") != "0" and 16 < DIM <= 32
                   else "bf16 three products"),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "backward": bwd,
        "index_match": {"vs": "oracle (oracle/vq_oracle.c)", "rows": sample_rows, "rate": match,
                        "vs_torch_port": {"what": "the reference's op sequence (model/vector_quantizer.py:88-119) as torch CPU "
                                                  "ops, matmul precision 'highest'", "rate": port_match,
                                          "mismatches": expl["mismatch"], "explained_by_fp32_near_tie": expl["explained"],
                                          "near_tie_explained_rate": (expl["explained"] / expl["mismatch"]) if expl["mismatch"] else 1.0}},
        "bulk_encode": bulk, "config1_forward": cfg1, "sweep_corners": corners,
        "check": {"loss": float(loss.item()), "perplexity": float(ppl.item()), "histogram_total": int(counts.sum().item())},
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--path", choices=["auto", "fma", "tc"], default="auto")
    ap.add_argument("--rows", type=int, default=N_VECTORS, help="vectors per GPU per step")
    ap.add_argument("--chunk-rows", type=int, default=1 << 19, help="host-path pipeline chunk (2^19 measured best: 357 M patches/s against 352 / 339 / 313 M for 2^20 / 2^21 / 2^22)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bwd", action="store_true")
    ap.add_argument("--no-bulk", action="store_true")
    ap.add_argument("--no-config1", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--bulk-chunk", type=int, default=65536, help="cycles per encode call of the bulk-encode object")
    ap.add_argument("--bulk-chunks", type=int, default=8, help="timed encode calls per GPU")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    # stdout carries the ONE JSON line and nothing else: libraries that write to fd 1 (NCCL prints its version banner
    # there) are pointed at stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global _emit
    def _emit(line: dict) -> None:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
