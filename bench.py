#!/usr/bin/env python
"""bench.py -- throughput of the AFEC low-level descriptor hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one pass of the hot path over one batch of synthetic decoded PCM per GPU.
Workloads (BASELINE.json `configs`):
  full     (default, configs[3] -- the north-star workload): ONE corpus of 12 500 x N mixed-length (0.5-30 s)
           44.1 kHz mono int16 files (100k files at 8 GPUs), sharded over the ranks by cost (afec_b200/shard.py),
           hop 1024, the full low-level descriptor set -- the same work `--impl reference` times on the host cores
  config2  configs[1]: 10 000 x 3-s one-shots per GPU, 2048-pt STFT hop 512 Hann + spectral stats
           (features = SPECTRAL); the reference has no such subset switch, so this line carries no cpu_baseline
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


def build_sharded_corpus(wl, rank, world):
    """ONE corpus of files_per_gpu x world files (every rank derives the same list from the same seeds); this rank's
    shard is chosen by cost = decoded length, longest-processing-time first (afec_b200/shard.py).
    -> (list of int16 arrays of this rank's shard, their indices in the global list)."""
    from afec_b200 import shard
    base = synth.corpus(N_UNIQUE, wl["seconds"], seed0=1000, min_seconds=wl["min_seconds"])
    n_total = wl["files_per_gpu"] * world
    costs = [len(base[i % N_UNIQUE]) for i in range(n_total)]
    mine = shard.my_shard(costs, rank, world) if world > 1 else list(range(n_total))
    return [base[i % N_UNIQUE] for i in mine], mine


def workload_config(wl, world):
    """The `config` object both arms print (identical by construction, so the two lines describe the same work)."""
    return {"workload": wl["desc"], "files_per_gpu": wl["files_per_gpu"], "files_total": wl["files_per_gpu"] * world,
            "hop": wl["hop"], "fft": 2048, "features": wl["features"], "unique_files": N_UNIQUE,
            "parallelism": "one corpus sharded by cost over the ranks (file batches), no collective",
            "l2": "every step's inputs (GBs of PCM, spectra) exceed the 126 MB L2; no flush needed"}


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
    """Bounded sample for the CPU legs: about 10 s of wall time per step on `cores` threads (the reference runs ~100 x real
    time per core on the GPU boxes' hosts at hop 1024: 4476 s of audio in 2.67 s on 16 threads, profiles/r02f_bench_full_n1.json;
    BASELINE.md section 2 quotes ~25 x on older hardware)."""
    avg_s = wl["seconds"] if wl["min_seconds"] is None else 0.5 * (wl["seconds"] + wl["min_seconds"])
    target_audio_s = 10.0 * 100.0 * cores
    return int(max(2 * cores, min(2048, target_audio_s / avg_s)))


def reference_sample(wl, cores):
    """The CPU legs' bounded sample of the workload: the first files of the global corpus (rank 0's view)."""
    n_sample = reference_sample_files(wl, cores)
    if wl["seconds"] >= 600:
        return build_corpus(dict(wl, files_per_gpu=min(n_sample, wl["files_per_gpu"])), 0)     # hour-long files: minutes of CPU time each
    if wl["features"] != "all":
        return build_corpus(dict(wl, files_per_gpu=n_sample), 0)
    return build_sharded_corpus(dict(wl, files_per_gpu=n_sample), 0, 1)[0]


def run_reference_arm(args, wl, rank, world):
    """The reference's own CPU implementation of the path (TSampleAnalyser::Extract into a TSqliteSampleDescriptorPool,
    Crawler.cpp:706-728) on every host thread, each step a bounded sample of the SAME workload (same corpus, hop and
    descriptor set as the GPU arm's `config`)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pcms = reference_sample(wl, cores)
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
        "dtype": "f64", "data": "synthetic", "config": workload_config(wl, max(1, args.gpus)),
        "same_feature_set": wl["features"] == "all",
        "cpu_baseline": {"value": value, "unit": "audio-hours/s", "cores": used, "kind": kind,
                         "sample": "%d files of the workload's corpus (%.1f s audio) per step, %d host threads, hop %d, full low-level "
                                   "set into a sqlite pool" % (len(pcms), audio / args.steps, used, wl["hop"])},
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
# algorithmic flops per frame and kernel group (rFFT counted 2.5 n log2 n); "rhythm" is per RHYTHM frame
GROUP_FLOPS = {"spectrum": 56320 + 6144 + 23600 + 8000, "peaks": 4000, "bands": 30000 + 2000 + 28672 + 392, "pitch": 184000,
               "autocorr": 280370, "rhythm": 29000, "stats": 0, "condition": 0}
GROUP_FLOPS_SPECTRAL_SUBSET = 78e3
FP32_NOMINAL_TFLOPS = 74.5      # the FP32 roofline the north_star target is stated against (2 x the FP64 pipe)
# what k_autocorr<2> (the FFT form, default) executes per frame: two 512-point complex FP32 transforms (5 n log2 n each), the
# bin-pair unpack / power / repack, the 33 aliased lags summed directly -- against the 280 k of the reference's direct sums
AUTOCORR_FFT_EXECUTED_FLOPS = 2 * 23040 + 10300 + 2200


def parity_check_sample(b, pcms, wl, n_check=32, seed=7):
    """After the timed region: `n_check` files of the batch that was just timed, compared with the oracle under the rules
    of tests/parity.py (the checker -- it never produces a result).  Returns (n_checked, mismatch strings)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    from afec_b200 import api
    from oracle import oracle
    oracle.build()
    rng = np.random.default_rng(seed)
    seen, errs, n = {}, [], 0
    order = rng.permutation(len(pcms)).tolist()
    only = None
    if wl["features"] == "spectral":
        only = ["spectral_rms", "spectral_centroid", "spectral_rolloff", "spectral_spread", "spectral_skewness",
                "spectral_kurtosis", "spectral_flatness", "spectral_flux"]
    for i in order:
        key = (pcms[i].ctypes.data, len(pcms[i]))
        if key in seen:
            continue                                  # tiled corpus: take distinct files
        seen[key] = True
        p = pcms[i]
        want = oracle.analyze(p, hop=wl["hop"], file_size=44 + 2 * p.size)
        data = oracle.condition(p)[0]
        e = parity.compare(b.result(i), want, only_series=only, check_stats=only is None, check_header=only is None,
                           mdata=data, hop=wl["hop"])
        errs += ["file %d: %s" % (i, x) for x in e[:3]]
        n += 1
        if n >= n_check:
            break
    return n, errs


def crawl_with_sink(pcms, wl, device, n_files=600):
    """The whole crawler on a bounded sample of the workload: WAV files on a RAM disk -> afec-b200-crawler (host decode
    ring, GPU, rows packed on the device, sqlite sink) -> afec-ll.db.  This is what a user's `Crawler -l low` run sees;
    the database writer, not the GPU, bounds it."""
    import shutil
    from afec_b200 import build as afx_build
    from oracle import oracle
    root = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    d = tempfile.mkdtemp(prefix="afx_crawl_", dir=root)
    try:
        audio_s = 0.0
        for i in range(min(n_files, len(pcms))):
            oracle.write_wav(os.path.join(d, "f%05d.wav" % i), pcms[i], wl.get("rate", 44100))
            audio_s += len(pcms[i]) / float(wl.get("rate", 44100))
        out = {}
        for tag, extra in (("one_writer", []), ("four_shards", ["--shards", "4"]), ("one_writer_32k_pages", ["--page-size", "32768"]),
                           ("one_writer_direct_load", ["--direct-load"]), ("four_shards_direct_load", ["--shards", "4", "--direct-load"])):
            try:
                db = os.path.join(d, tag + ".db")
                t0 = time.perf_counter()
                r = subprocess.run([afx_build.CRAWLER, "-o", db, "--hop", str(wl["hop"]), "--devices", str(device), "-j", "3"] + extra + [d],
                                   capture_output=True, text=True, timeout=600)
                wall = time.perf_counter() - t0
                js = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
                size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.startswith(tag + ".db"))
                out[tag] = {"value": js["audio_seconds"] / 3600.0 / js["seconds"], "unit": "audio-hours/s", "files": js["files"],
                            "seconds": js["seconds"], "wall_seconds": wall, "rows_per_s": js["files"] / js["seconds"],
                            "db_mb_per_s": size / js["seconds"] / 1e6}
                for f in os.listdir(d):                 # the next variant starts with the RAM disk as this one found it
                    if f.startswith(tag + ".db"):
                        os.remove(os.path.join(d, f))
            except Exception as e:                      # one variant failing must not take the others (or the bench line) with it
                out[tag] = {"value": 0.0, "unit": "audio-hours/s", "error": repr(e)[:200]}
        best = max(out, key=lambda k: out[k]["value"])
        return dict(out[best], writer=best, variants=out, sample="%d WAV files of the workload's corpus (%.0f s audio) on %s" % (min(n_files, len(pcms)), audio_s, root),
                    path="afec-b200-crawler: decode threads -> pinned ring -> GPU (rows packed on the device) -> afec-ll.db (variants: journal-less "
                         "bulk load through sqlite with one / four writers, 32 KB pages, the file written directly in sqlite's format)")
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)      # 12 x 0.49 s: a timed region above 5 s on the default workload
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="full", choices=sorted(WORKLOADS))
    ap.add_argument("--seconds", type=float, default=0.0, help="override file duration (debug)")
    ap.add_argument("--files", type=int, default=0, help="override files per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-sink", action="store_true", help="skip the crawler-with-sqlite-sink leg")
    ap.add_argument("--parts", type=int, default=0, help="long workload at N = 1: condition every file in this many parts")
    ap.add_argument("--e2e-slots", type=int, default=3)
    ap.add_argument("--e2e-chunks", type=int, default=12)
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

    import ctypes as C
    import itertools
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

    def reduce(x: float, op) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX) if dist is not None else x

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM) if dist is not None else x

    feats = api.FEAT_SPECTRAL if wl["features"] == "spectral" else api.FEAT_ALL
    an = api.SampleAnalyser(44100, 2048, wl["hop"], device=local_rank, features=feats)

    # ---- synthetic decoded PCM in ONE pinned arena (what per-GPU decode threads would fill) ----
    if wl["features"] == "all" and wl["seconds"] < 600:
        pcms, _ = build_sharded_corpus(wl, rank, world)
    else:
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

    def wrap(sl, arr, n):
        h = C.c_void_p()
        sl._check(sl._L.afx_batch_create(sl._ctx, arr, n, C.byref(h)))
        b = api.Batch.__new__(api.Batch)
        b._an, b._L, b._keep, b.n_files, b._files, b._h = sl, sl._L, None, n, arr, h
        return b

    # ---- device-resident timing: inputs already in HBM, K passes of the kernels -------------------
    b = wrap(an, files_arr, len(files))
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
    dev_ms_max = max_over_ranks(dev_ms)
    total_audio_hours = sum_over_ranks(audio_hours)
    total_frames = sum_over_ranks(float(cnt["main_frames"]))
    total_rframes = sum_over_ranks(float(cnt["rhythm_frames"]))
    value = total_audio_hours * args.steps / (dev_ms_max / 1000.0)
    frames_per_s = total_frames * args.steps / (dev_ms_max / 1000.0)

    # ---- what was timed is what the reference computes: sampled files of THIS batch against the oracle ----
    parity_n, parity_errs = 0, []
    if rank == 0 and not args.no_parity_check:
        try:
            parity_n, parity_errs = parity_check_sample(b, pcms, wl)
        except Exception as e:
            parity_errs = ["parity check failed to run: %r" % (e,)]
    b.free()

    # ---- end to end through the C ABI: host PCM -> H2D -> kernels -> D2H results, every step ------
    # Host pipeline as in afec_b200/host/gpu_analyser.cpp: slots = contexts (stream + device buffers each), one
    # host thread per slot; a step's files are cut into chunks that the slot threads claim in turn, so the H2D /
    # D2H copies of one chunk overlap the kernels of another.  ctypes releases the GIL inside the C calls.
    E2E_SLOTS, E2E_CHUNKS = max(1, args.e2e_slots), max(1, min(args.e2e_chunks, len(files)))
    slots = [an] + [api.SampleAnalyser(44100, 2048, wl["hop"], device=local_rank, features=feats) for _ in range(E2E_SLOTS - 1)]
    bounds = [len(files) * i // E2E_CHUNKS for i in range(E2E_CHUNKS + 1)]
    chunk_arrs = [(api.AfxFile * (bounds[i + 1] - bounds[i]))(*files[bounds[i]:bounds[i + 1]]) for i in range(E2E_CHUNKS)]
    e2e_bytes = [0, 0]

    def pipeline_steps(n_steps: int, copy_only: bool):
        """n_steps passes over the corpus as ONE stream of chunks through the slots (a crawler's file list does not stop
        between passes either: the pipeline fills once and drains once per call); every chunk is uploaded from pinned host
        memory and its results are read back on the host."""
        counter = itertools.count()
        lock = threading.Lock()
        tot = [0, 0, 0.0]

        def work(sl):
            while True:
                with lock:
                    ci = next(counter)
                if ci >= E2E_CHUNKS * n_steps:
                    return
                arr = chunk_arrs[ci % E2E_CHUNKS]
                bb = wrap(sl, arr, len(arr))
                if copy_only:
                    bb.upload(); bb.sync()
                    c = bb.counters()
                    with lock:
                        tot[0] += c["h2d_bytes"]
                else:
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
        e2e_bytes[0], e2e_bytes[1] = tot[0] // n_steps, tot[1] // n_steps

    pipeline_steps(2, False)
    barrier()
    t0 = time.perf_counter()
    pipeline_steps(args.steps, False)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    h2d, d2h = e2e_bytes
    e2e_value = total_audio_hours * args.steps / e2e_s
    # copy-only leg: the same arena, chunks, slots and threads, uploads only -- what the host -> device fabric gives
    # every rank while all ranks copy at once (separates a PCIe / host-memory ceiling from the kernels)
    pipeline_steps(1, True)
    barrier()
    t0 = time.perf_counter()
    n_copy = max(2, args.steps // 2)
    pipeline_steps(n_copy, True)
    torch.cuda.synchronize()
    copy_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    h2d_gbs = e2e_bytes[0] * n_copy / copy_s / 1e9
    for sl in slots[1:]:
        sl.close()
    an.trim()          # the roofline leg below runs the same batch on a second context: give this one's device buffers back

    # ---- roofline: per-kernel-group device times (CUDA events around each group, one stream), timed live -----------
    roof = None
    peaks, peak_src = measured_peaks()
    try:
        os.environ["AFX_DEBUG_KERNEL_TIMES"] = "1"
        an2 = api.SampleAnalyser(44100, 2048, wl["hop"], device=local_rank, features=feats)
        del os.environ["AFX_DEBUG_KERNEL_TIMES"]
        b2 = wrap(an2, files_arr, len(files))
        b2.upload()
        for _ in range(3):
            b2.compute()
        b2.sync()
        acc = {}
        n_roof = max(2, min(args.steps, 5))
        for _ in range(n_roof):
            b2.compute(); b2.sync()
            for name, ms in b2.kernel_times():
                acc[name] = acc.get(name, 0.0) + ms
        b2.download(); b2.sync()
        c2 = b2.counters()
        frames, rframes = c2["main_frames"], c2["rhythm_frames"]
        fp64_peak = an2.fp64_peak_tflops()
        b2.free(); an2.close()
        groups = {k: v / n_roof for k, v in acc.items()}
        top = max(groups, key=groups.get)
        top_ms = groups[top]
        H = wl["hop"]
        if wl["features"] == "spectral":
            bytes_per_frame = 2 * H + 8 * 8          # int16 hop in, 8 float64 descriptors out
            gflops = {"spectrum": GROUP_FLOPS_SPECTRAL_SUBSET * frames}
        else:
            bytes_per_frame = 2 * H + 136 * 8 + 2 * 8 * (H // 128)
            gflops = {g: (f * rframes if g == "rhythm" else f * frames) for g, f in GROUP_FLOPS.items()}
        step_flops = sum(gflops.values())
        table = {g: {"ms": ms, "share": ms / sum(groups.values()), "algorithmic_gflop": gflops.get(g, 0.0) / 1e9,
                     "tflops": gflops.get(g, 0.0) / (ms * 1e-3) / 1e12 if ms > 0 else None,
                     "frac_fp64": gflops.get(g, 0.0) / (ms * 1e-3) / 1e12 / fp64_peak if ms > 0 and fp64_peak else None}
                 for g, ms in groups.items()}
        # the autocorrelation multiplies in FP32 since round 2 (unless AFX_AUTOCORR_FP64=1): its pipe is the FP32 one
        if "autocorr" in table and table["autocorr"]["tflops"] is not None and os.environ.get("AFX_AUTOCORR_FP64", "0") in ("", "0"):
            table["autocorr"]["pipe"] = "fp32"
            if os.environ.get("AFX_AUTOCORR_DIRECT", "0") in ("", "0"):
                ex = AUTOCORR_FFT_EXECUTED_FLOPS * frames
                table["autocorr"]["form"] = ("FFT form: R = IFFT(|FFT(x)|^2) on one warp per frame; algorithmic_gflop / tflops / frac_fp64 count the "
                                             "reference's direct sums (SURVEY.md 8(d)), executed_* what the kernel issues")
                table["autocorr"]["executed_gflop"] = ex / 1e9
                table["autocorr"]["executed_tflops"] = ex / (table["autocorr"]["ms"] * 1e-3) / 1e12
                table["autocorr"]["frac_fp32_nominal"] = table["autocorr"]["executed_tflops"] / FP32_NOMINAL_TFLOPS
            else:
                table["autocorr"]["frac_fp32_nominal"] = table["autocorr"]["tflops"] / FP32_NOMINAL_TFLOPS
        top_tf = gflops.get(top, 0.0) / (top_ms * 1e-3) / 1e12
        step_ms = dev_ms / args.steps                     # this rank's step (the `value` timing)
        step_tf = step_flops / (step_ms * 1e-3) / 1e12
        # the same with the autocorrelation counted at what its FFT form executes instead of the reference's direct sums
        step_flops_exec = step_flops - (gflops.get("autocorr", 0.0) - table.get("autocorr", {}).get("executed_gflop", gflops.get("autocorr", 0.0) / 1e9) * 1e9)
        step_tf_exec = step_flops_exec / (step_ms * 1e-3) / 1e12
        roof = {
            "bound": "fp64", "achieved": top_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": top_tf / fp64_peak if fp64_peak else None,
            "traffic": None, "kernel": top, "kernel_ms": top_ms, "kernel_share_of_step": top_ms / sum(groups.values()),
            "peak_source": "measured live: dependent-free DFMA loop (afx_measure_fp64_peak); MEASURED_PEAKS.json holds no FP64 figure",
            "algorithmic_flops": gflops.get(top, 0.0), "frames_per_launch": frames, "rhythm_frames_per_launch": rframes,
            "note": "the frame kernels compute in FP64 (bit-for-bit decisions of the reference: peak counts, onset thresholds) except the "
                    "autocorrelation (FP32, FFT form, see groups.autocorr); achieved = SURVEY.md 8(d) algorithmic flops of the group / "
                    "its CUDA-event time",
            "step": {"algorithmic_flops": step_flops, "ms": step_ms, "achieved": step_tf, "unit": "TFLOP/s",
                     "frac_fp64": step_tf / fp64_peak if fp64_peak else None, "frac_fp32_nominal": step_tf / FP32_NOMINAL_TFLOPS,
                     "flops_per_main_frame": step_flops / max(1, frames),
                     "executed_flops": step_flops_exec, "executed_tflops": step_tf_exec,
                     "frac_fp64_executed": step_tf_exec / fp64_peak if fp64_peak else None,
                     "executed_note": "autocorr counted at the flops of its FFT form; every other group executes the reference's formulation"},
            "hbm": {"achieved": bytes_per_frame * frames / (step_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": bytes_per_frame * frames / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "algorithmic_bytes_per_frame": bytes_per_frame, "peak_source": peak_src + " (MEASURED_PEAKS.json)"},
            "groups": table,
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
        roof = {"bound": "fp64", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None,
                "traffic": None, "error": repr(e)}

    # ---- reported CPU baseline (rank 0, N = 1 only, bounded sample of the same workload) -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and wl["features"] == "all":
        try:
            cores = os.cpu_count() or 1
            sample = reference_sample(wl, cores)
            r = cpu_reference_run(sample, wl["hop"], cores, rate=wl.get("rate", 44100))
            cpu = {"value": r["audio_hours_per_s"], "unit": "audio-hours/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": "%d files of the workload's corpus (%.1f s audio), hop %d, %d host threads, full low-level set into a sqlite pool"
                             % (len(sample), r["audio_s"], wl["hop"], r["cores"]), "seconds": r["seconds"]}
        except Exception as e:
            cpu = {"value": None, "unit": "audio-hours/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}

    sink = None
    if rank == 0 and world == 1 and wl["features"] == "all" and not args.no_sink:
        try:
            sink = crawl_with_sink(pcms, wl, local_rank)
        except Exception as e:
            sink = {"value": None, "error": repr(e)}

    if rank == 0:
        line = {
            "metric": "low-level descriptor throughput (audio-hours/sec)", "value": value, "unit": "audio-hours/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl, world),
            "frames_per_s": frames_per_s, "rhythm_frames_per_s": total_rframes * args.steps / (dev_ms_max / 1000.0),
            "main_frames_per_step": total_frames,
            "e2e": {"value": e2e_value, "unit": "audio-hours/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": 1000.0 * e2e_s / args.steps, "h2d_gbs_per_gpu": h2d_gbs,
                    "h2d_note": "copy-only leg: same pinned arena / chunks / slot threads with no kernels, all ranks at once, max over ranks",
                    "path": "afx_batch_create -> upload (pinned H2D) -> compute -> download (D2H) -> sync per chunk; %d chunks per step over "
                            "%d contexts / host threads (copies overlap kernels; the K timed steps are one continuous stream of chunks, the "
                            "computes of the contexts run one after the other on the device)" % (E2E_CHUNKS, E2E_SLOTS)},
            "gpu_launches": int(cnt["kernel_launches"]) * args.steps * world,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "parity_checked": parity_n, "parity_mismatches": parity_errs[:8], "e2e_with_sink": sink,
        }
        print(json.dumps(line), flush=True)
    arena.free()
    an.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
