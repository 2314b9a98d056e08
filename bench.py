#!/usr/bin/env python
"""bench.py -- throughput of the AFEC low-level descriptor hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one pass of the hot path over one batch of synthetic decoded PCM per GPU.
Workloads (BASELINE.json `configs`):
  config2  (default at N=1, configs[1]): 10 000 x 3-s 44.1 kHz mono int16 one-shots per GPU,
           2048-pt STFT hop 512 Hann + spectral stats (features = SPECTRAL)
  full     configs[3] per-GPU shard: 12 500 mixed-length (0.5-30 s) files per GPU, hop 1024, the
           full low-level descriptor set (100k files over 8 GPUs)
Multi-GPU: files are independent units, each rank owns its own shard and its own context; there is
no data-path collective.  torch.distributed (NCCL) is used only for the barrier and the max-over-ranks
reduction of the timings.  `value` = audio-hours of all ranks / max-over-ranks device time.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/afec_ref, the unmodified
AFEC sources compiled by oracle/build_ref.sh; else the C port oracle/libafec_oracle.so) on the host
cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from afec_b200 import synth  # noqa: E402

WORKLOADS = {
    "config2": dict(files_per_gpu=10000, seconds=3.0, min_seconds=None, hop=512, features="spectral",
                    desc="configs[1]: 2048-pt STFT hop 512 Hann + spectral stats (centroid/flatness/rolloff/flux/RMS) "
                         "on 10k 3-s 44.1 kHz mono int16 one-shots per GPU"),
    "full": dict(files_per_gpu=12500, seconds=30.0, min_seconds=0.5, hop=1024, features="all",
                 desc="configs[3] shard: full low-level descriptor set on 12.5k mixed-length (0.5-30 s) 44.1 kHz mono "
                      "int16 files per GPU (100k files at 8 GPUs)"),
}
WORKLOADS["long"] = dict(files_per_gpu=2, seconds=3600.0, min_seconds=None, hop=1024, features="all", rate=96000, channels=2,
                         desc="configs[4] on one GPU: full low-level set on 1-hour 96 kHz stereo int16 files (fused downmix + "
                              "libresample-exact 96k->44.1k resample dominate; analysis is capped at 20 s by the reference)")
N_UNIQUE = 64            # distinct synthetic files, tiled to the workload size


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_corpus(wl, rank):
    """-> list of int16 arrays (tiled)."""
    if wl.get("rate", 44100) != 44100 or wl.get("channels", 1) != 1:
        # long multi-channel files: a 30-s clip repeated to the requested duration
        clip = synth.one_shot(7 + rank, 30.0, rate=wl["rate"], channels=wl["channels"])
        reps = max(1, int(round(wl["seconds"] / 30.0)))
        one = np.ascontiguousarray(np.tile(clip, (reps, 1)))
        return [one for _ in range(wl["files_per_gpu"])]
    return synth.tiled_corpus(wl["files_per_gpu"], N_UNIQUE, seconds=wl["seconds"], seed0=1000 * rank,
                              min_seconds=wl["min_seconds"])


# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(files_pcm, hop, threads, reps=1, rate=44100):
    """Times the reference (or the port) on host cores over the given files.  Returns dict."""
    from oracle import oracle
    audio_s = sum(len(p) for p in files_pcm) / float(rate) * reps
    if oracle.have_reference():
        with tempfile.TemporaryDirectory() as d:
            paths = []
            for i, p in enumerate(files_pcm):
                path = os.path.join(d, "f%05d.wav" % i)
                oracle.write_wav(path, p, rate)
                paths.append(path)
            env = dict(os.environ, HOME=d)
            t0 = time.perf_counter()
            out = subprocess.run([oracle.REF_BIN, "bench", str(hop), str(threads), str(reps), os.path.join(d, "ll.db")] + paths,
                                 check=True, env=env, capture_output=True, text=True).stdout
            wall = time.perf_counter() - t0
            js = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
            secs = js["seconds"]
        kind = "reference"
        del wall
    else:
        oracle.build()
        t0 = time.perf_counter()
        for _ in range(reps):
            for p in files_pcm:
                oracle.analyze(p, hop=hop, src_rate=rate)
        secs = time.perf_counter() - t0
        kind, threads = "port", 1
    return dict(seconds=secs, audio_hours_per_s=audio_s / 3600.0 / secs, kind=kind, cores=threads, audio_s=audio_s)


def reference_sample_files(wl, cores):
    """Bounded sample for the CPU legs: about 10 s of wall time per step on `cores` threads
    (the reference runs ~25 x real time per core)."""
    avg_s = wl["seconds"] if wl["min_seconds"] is None else 0.5 * (wl["seconds"] + wl["min_seconds"])
    target_audio_s = 10.0 * 25.0 * cores
    return int(max(2 * cores, min(2048, target_audio_s / avg_s)))


def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = reference_sample_files(wl, cores)
    if wl["seconds"] >= 600:
        n_sample = min(n_sample, wl["files_per_gpu"])          # hour-long files: minutes of CPU time each
    pcms = build_corpus(dict(wl, files_per_gpu=n_sample), 0)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_run(pcms[: max(4, cores // 2)], wl["hop"], cores, rate=wl.get("rate", 44100))
    secs, audio = 0.0, 0.0
    kind = "port"
    for _ in range(args.steps):
        r = cpu_reference_run(pcms, wl["hop"], cores, rate=wl.get("rate", 44100))
        secs += r["seconds"]; audio += r["audio_s"]; kind = r["kind"]; used = r["cores"]
    value = audio / 3600.0 / secs
    line = {
        "impl": "reference", "metric": "low-level descriptor throughput (audio-hours/sec)", "value": value,
        "unit": "audio-hours/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "hop": wl["hop"], "note": "reference CPU path computes the FULL low-level set "
                   "(TSampleAnalyser::Extract into a sqlite pool, as Crawler.cpp:706-728) whatever the workload's subset"},
        "cpu_baseline": {"value": value, "unit": "audio-hours/s", "cores": used, "kind": kind,
                         "sample": "%d files (%.1f s audio) per step, %d host threads" % (len(pcms), audio / args.steps, used)},
        "e2e": {"value": value, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_long_sharded(args, wl, rank, local_rank, world):
    """BASELINE configs[4]: every long file is cut into sample-range parts, one per rank (afec_b200/longfile.py):
    each rank uploads and conditions ONLY its slice; three all-gathers of a 112-byte record combine the per-file
    reductions and one reduce assembles the <= 20 s analysis window on the file's analysis rank.  Total work is
    fixed (strong scaling).  Every step moves the PCM host -> device, so `value` is an end-to-end number."""
    import torch
    import torch.distributed as dist
    from afec_b200 import api, longfile
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_files = args.files or 10
    rate, nch = wl["rate"], wl["channels"]
    clip = synth.one_shot(7, 30.0, rate=rate, channels=nch)
    nframes = clip.shape[0] * max(1, int(round(wl["seconds"] / 30.0)))
    an = api.SampleAnalyser(44100, 2048, wl["hop"], device=local_rank, features=api.FEAT_ALL)
    n_parts = world if world > 1 else max(1, args.parts)
    parts = longfile.plan_parts(nframes, rate, n_parts)
    whole = longfile.describe_whole(nframes, nch, rate, np.int16)
    # the slices this rank will need (file k gives part (rank - k) mod world to this rank), in pinned memory
    my_parts = sorted({(rank - k) % n_parts for k in range(n_files)}) if world > 1 else list(range(n_parts))
    arenas, slices = {}, {}
    for p in my_parts:
        sb, se = parts[p][0], parts[p][1]
        a = an.pinned(max(1, (se - sb) * nch * 2))
        v = a.array.view(np.int16)[: (se - sb) * nch].reshape(se - sb, nch)
        pos = sb
        while pos < se:                                     # the file is the 30-s clip repeated
            o = pos % clip.shape[0]
            m = min(se - pos, clip.shape[0] - o)
            v[pos - sb:pos - sb + m] = clip[o:o + m]
            pos += m
        arenas[p], slices[p] = a, v

    def step():
        frames = 0
        nxt = None
        for k in range(n_files):
            if world > 1:
                # the next file's part is opened (asynchronous H2D) before this file's host-mediated phases
                p = (rank - k) % world
                job = nxt if nxt is not None else longfile.PartJob(an, whole, parts[p], slices[p])
                nxt = None
                if k + 1 < n_files:
                    p1 = (rank - k - 1) % world
                    nxt = longfile.PartJob(an, whole, parts[p1], slices[p1])
                b = longfile.analyze_sharded(an, whole, parts[p], slices[p], dist, analysis_rank=k % world, device=dev, job=job)
            else:
                jobs = [longfile.PartJob(an, whole, parts[p], slices[p]) for p in range(n_parts)]
                g = longfile.merge_sums([j.peak() for j in jobs])
                g = longfile.merge_sums([j.trim(g) for j in jobs])
                g = longfile.merge_sums([j.effective(g) for j in jobs])
                begin, count = longfile.window_of(an, whole, g)
                win = np.zeros(count, dtype=np.float32)
                for j in jobs:
                    j.read(begin, count, win); j.close()
                b = longfile.analyze_conditioned(an, whole, g, win, begin)
            if b is not None:
                frames += b.raw_result(0).n_frames
                b.free()
        return frames

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        frames += step()
    barrier()
    secs = time.perf_counter() - t0
    clocks = sampler.stop()
    t = torch.tensor([secs, float(frames)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        secs, frames = float(tm[0]), float(ts[1])
    audio_hours = n_files * nframes / float(rate) / 3600.0
    value = audio_hours * args.steps / secs
    h2d = sum((parts[p][1] - parts[p][0]) * nch * 2 for p in range(n_parts)) * n_files
    if rank == 0:
        peaks, peak_src = measured_peaks()
        line = {
            "metric": "low-level descriptor throughput (audio-hours/sec)", "value": value, "unit": "audio-hours/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": 1000.0 * secs / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[4]: %d x 1-hour 96 kHz stereo int16 files, each cut into %d sample-range parts (one per GPU): "
                                   "fused downmix + libresample-exact 96k->44.1k resample + peak/RMS/trim per part, host-combined "
                                   "reductions, full low-level set on the 20 s the reference analyses" % (n_files, n_parts),
                       "files": n_files, "parts": n_parts, "hop": wl["hop"], "fft": 2048, "features": "all",
                       "parallelism": "sample-range parts across ranks; 3 all-gathers of 112 B + 1 window reduce per file",
                       "l2": "every step re-uploads the PCM (%.2f GB per file) from pinned host memory" % (nframes * nch * 2 / 1e9),
                       "timing": "host clock between barrier + cudaDeviceSynchronize pairs, max over ranks (the phases are host-mediated)"},
            "frames_per_s": frames / secs, "main_frames_per_step": frames / args.steps,
            "e2e": {"value": value, "unit": "audio-hours/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": None,
                    "ms_per_step": 1000.0 * secs / args.steps, "path": "afx_part_open/peak/trim/effective/read -> afx_analyze_conditioned"},
            "gpu_launches": None, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": h2d * args.steps / secs / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": h2d * args.steps / secs / 1e9 / peaks["hbm_gbs"], "traffic": None,
                         "note": "PCIe-bound: the PCM crosses host -> device once per step; achieved = PCM bytes / step time"},
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    for a in arenas.values():
        a.free()
    an.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--seconds", type=float, default=0.0, help="override file duration (debug)")
    ap.add_argument("--files", type=int, default=0, help="override files per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parts", type=int, default=0, help="long workload at N = 1: condition every file in this many parts")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.files:
        wl["files_per_gpu"] = args.files
    if args.seconds:
        wl["seconds"] = args.seconds

    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return
    if args.workload == "long" and (world > 1 or args.parts):
        run_long_sharded(args, wl, rank, local_rank, world)
        return

    import torch
    from afec_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the descriptor path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    feats = api.FEAT_SPECTRAL if wl["features"] == "spectral" else api.FEAT_ALL
    an = api.SampleAnalyser(44100, 2048, wl["hop"], device=local_rank, features=feats)

    # ---- synthetic decoded PCM in ONE pinned arena (what per-GPU decode threads would fill) ----
    pcms = build_corpus(wl, rank)
    rate, nch = wl.get("rate", 44100), wl.get("channels", 1)
    total = sum(p.size for p in pcms)
    arena = an.pinned(total * 2)
    files, off = [], 0
    view = arena.array.view(np.int16)
    for p in pcms:
        view[off:off + p.size] = p.reshape(-1)
        files.append(api.AfxFile(arena.ptr + 2 * off, p.shape[0], nch, rate, api.AFX_PCM_I16, 16, 44 + 2 * p.size))
        off += p.size
    audio_hours = total / float(nch) / float(rate) / 3600.0
    files_arr = (api.AfxFile * len(files))(*files)

    def make_batch():
        import ctypes as C
        h = C.c_void_p()
        an._check(an._L.afx_batch_create(an._ctx, files_arr, len(files), C.byref(h)))
        b = api.Batch.__new__(api.Batch)
        b._an, b._L, b._keep, b.n_files, b._files, b._h = an, an._L, None, len(files), files_arr, h
        return b

    # ---- device-resident timing: inputs already in HBM, K passes of the kernels -------------------
    b = make_batch()
    b.upload(); b.sync()
    for _ in range(max(3, args.warmup)):
        b.compute()
    b.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.perf_counter()
    while not sampler.rows and time.perf_counter() - t_wait < 3.0:     # nvidia-smi needs ~1 s to deliver its first row:
        b.compute(); b.sync()                                          # keep the GPU under the same load meanwhile
    barrier()
    dev_ms = 0.0
    for _ in range(args.steps):
        b.compute(); b.sync()
        dev_ms += b.timings()[1]
    barrier()
    for _ in range(3):                                                 # a short timed region may end between two samples
        if len(sampler.rows) >= 2:
            break
        b.compute(); b.sync()
    clocks = sampler.stop()
    b.download(); b.sync()
    cnt = b.counters()
    ktimes = b.kernel_times()
    dev_ms_max = max_over_ranks(dev_ms)
    total_audio_hours = sum_over_ranks(audio_hours)
    total_frames = sum_over_ranks(float(cnt["main_frames"]))
    value = total_audio_hours * args.steps / (dev_ms_max / 1000.0)
    frames_per_s = total_frames * args.steps / (dev_ms_max / 1000.0)
    b.free()

    # ---- end to end through the C ABI: host PCM -> H2D -> kernels -> D2H results, every step ------
    # Host pipeline as in afec_b200/host/gpu_analyser.cpp: E2E_SLOTS contexts (stream + device buffers each), one
    # host thread per slot; a step's files are cut into chunks that the slot threads claim in turn, so the H2D /
    # D2H copies of one chunk overlap the kernels of another.  ctypes releases the GIL inside the C calls.
    E2E_SLOTS, E2E_CHUNKS = 3, min(12, len(files))
    slots = [an] + [api.SampleAnalyser(44100, 2048, wl["hop"], device=local_rank, features=feats) for _ in range(E2E_SLOTS - 1)]
    bounds = [len(files) * i // E2E_CHUNKS for i in range(E2E_CHUNKS + 1)]
    chunk_arrs = [(api.AfxFile * (bounds[i + 1] - bounds[i]))(*files[bounds[i]:bounds[i + 1]]) for i in range(E2E_CHUNKS)]
    import ctypes as C
    import itertools
    e2e_bytes = [0, 0]

    def e2e_step():
        counter = itertools.count()
        lock = threading.Lock()
        tot = [0, 0, 0.0]

        def work(sl):
            while True:
                with lock:
                    ci = next(counter)
                if ci >= E2E_CHUNKS:
                    return
                arr = chunk_arrs[ci]
                h = C.c_void_p()
                sl._check(sl._L.afx_batch_create(sl._ctx, arr, len(arr), C.byref(h)))
                bb = api.Batch.__new__(api.Batch)
                bb._an, bb._L, bb._keep, bb.n_files, bb._files, bb._h = sl, sl._L, None, len(arr), arr, h
                bb.run()
                r = bb.raw_result(len(arr) - 1)            # the step's result is read on the host
                c = bb.counters()
                with lock:
                    tot[0] += c["h2d_bytes"]; tot[1] += c["d2h_bytes"]; tot[2] += r.header[1]
                bb.free()
        ths = [threading.Thread(target=work, args=(sl,)) for sl in slots]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        e2e_bytes[0], e2e_bytes[1] = tot[0], tot[1]

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    h2d, d2h = e2e_bytes
    e2e_value = total_audio_hours * args.steps / e2e_s
    for sl in slots[1:]:
        sl.close()
    an.trim()          # the roofline leg below runs the same batch on a second context: give this one's device buffers back

    # ---- roofline of the dominant kernel (k_spectrum), timed live with CUDA events ---------------
    roof = None
    peaks, peak_src = measured_peaks()
    try:
        os.environ["AFX_DEBUG_KERNEL_TIMES"] = "1"
        an2 = api.SampleAnalyser(44100, 2048, wl["hop"], device=local_rank, features=feats)
        import ctypes as C
        h = C.c_void_p()
        an2._check(an2._L.afx_batch_create(an2._ctx, files_arr, len(files), C.byref(h)))
        b2 = api.Batch.__new__(api.Batch)
        b2._an, b2._L, b2._keep, b2.n_files, b2._files, b2._h = an2, an2._L, None, len(files), files_arr, h
        b2.upload()
        for _ in range(3):
            b2.compute()
        b2.sync()
        acc = {}
        for _ in range(args.steps):
            b2.compute(); b2.sync()
            for name, ms in b2.kernel_times():
                acc[name] = acc.get(name, 0.0) + ms
        b2.download(); b2.sync()
        frames = b2.counters()["main_frames"]
        fp64_peak = an2.fp64_peak_tflops()
        b2.free(); an2.close()
        del os.environ["AFX_DEBUG_KERNEL_TIMES"]
        groups = {k: v / args.steps for k, v in acc.items()}
        top = max(groups, key=groups.get)
        top_ms = groups[top]
        # algorithmic bytes / flops per main frame (DESIGN.md "Kernels")
        H = wl["hop"]
        if wl["features"] == "spectral":
            bytes_per_frame = 2 * H + 8 * 8          # int16 hop in, 8 float64 descriptors out
            flops_per_frame = 78e3
        else:
            bytes_per_frame = 2 * H + 136 * 8 + 2 * 8 * (H // 128)
            flops_per_frame = 0.855e6
        share = top_ms / sum(groups.values())
        hbm_achieved = bytes_per_frame * frames / (top_ms * 1e-3) / 1e9
        fl_achieved = flops_per_frame * frames * share / (top_ms * 1e-3) / 1e12 if wl["features"] != "spectral" else \
            flops_per_frame * frames / (top_ms * 1e-3) / 1e12
        roof = {
            "bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": hbm_achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src + " (MEASURED_PEAKS.json)",
            "kernel": top, "kernel_ms": top_ms, "kernel_share_of_step": share,
            "algorithmic_bytes_per_frame": bytes_per_frame, "frames_per_launch": frames,
            "fp64": {"bound": "fp64 fma pipe (the path is compute-bound: ~%d flop/B)" % int(flops_per_frame / bytes_per_frame),
                     "achieved": fl_achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": fl_achieved / fp64_peak if fp64_peak else None,
                     "peak_source": "measured live: dependent-free DFMA loop (afx_measure_fp64_peak)",
                     "algorithmic_flops_per_frame": flops_per_frame},
            "groups_ms": groups,
        }
        # DRAM traffic of the dominant kernel group from the committed `ncu --set full` capture (bytes per main frame
        # there x the frames of this launch); the capture is of the same kernels on a smaller batch of the same files
        prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(prof):
            with open(prof) as f:
                tr = json.load(f).get(wl["features"], {})
            per_frame = tr.get("groups", {}).get(top, {}).get("dram_bytes_per_main_frame")
            if per_frame and tr.get("hop") == wl["hop"]:
                roof["traffic"] = per_frame * frames
                roof["traffic_source"] = tr.get("source")
    except Exception as e:  # the roofline leg must not take the headline number down
        roof = {"bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None,
                "traffic": None, "error": repr(e)}

    # ---- reported CPU baseline (rank 0, N = 1 only, bounded sample) -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            sample = pcms[: reference_sample_files(wl, cores)]
            r = cpu_reference_run(sample, wl["hop"], cores, rate=wl.get("rate", 44100))
            cpu = {"value": r["audio_hours_per_s"], "unit": "audio-hours/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": "%d files (%.1f s audio), hop %d, %d host threads, FULL low-level set into a sqlite pool"
                             % (len(sample), r["audio_s"], wl["hop"], r["cores"]), "seconds": r["seconds"]}
        except Exception as e:
            cpu = {"value": None, "unit": "audio-hours/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}

    if rank == 0:
        line = {
            "metric": "low-level descriptor throughput (audio-hours/sec)", "value": value, "unit": "audio-hours/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "files_per_gpu": wl["files_per_gpu"], "hop": wl["hop"], "fft": 2048,
                       "features": wl["features"], "parallelism": "file-batch shard per GPU, no collective",
                       "l2": "inputs (%.2f GB PCM + spectra per step) exceed the 126 MB L2; no flush needed" % (total * 2 / 1e9),
                       "unique_files": N_UNIQUE},
            "frames_per_s": frames_per_s, "main_frames_per_step": total_frames,
            "e2e": {"value": e2e_value, "unit": "audio-hours/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": 1000.0 * e2e_s / args.steps,
                    "path": "afx_batch_create -> upload (pinned H2D) -> compute -> download (D2H) -> sync per chunk; 12 chunks per step over 3 contexts / host threads (copies overlap kernels)"},
            "gpu_launches": int(cnt["kernel_launches"]) * args.steps * world,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "kernel_group_ms": dict(ktimes) if ktimes else None,
        }
        print(json.dumps(line), flush=True)
    arena.free()
    an.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
