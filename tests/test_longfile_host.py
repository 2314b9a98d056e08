"""Host logic of the long-file path (BASELINE config 5) on CPU: part planning (pure host arithmetic of the C ABI),
the combine of the per-part reductions, and the multi-process protocol over gloo (world_size 2) with a numpy
part worker standing in for the GPU one.  The numpy worker restates SampleAnalyser.cpp:612-701, 1715-1756 for
an already-mono 44.1 kHz signal; its combined result is checked against the C oracle's conditioning."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from afec_b200 import api, longfile, synth


@pytest.mark.parametrize("nframes,rate,n_parts", [(96000 * 60, 96000, 4), (96000 * 60, 96000, 8), (44100 * 100, 44100, 3),
                                                   (22050 * 7, 22050, 8), (48000 * 3, 48000, 2), (5000, 96000, 4), (44100, 44100, 1)])
def test_plan_covers_file_and_carries_halo(nframes, rate, n_parts):
    parts = longfile.plan_parts(nframes, rate, n_parts)
    n = nframes if rate == 44100 else max(1, int(nframes / (rate / 44100.0) + 0.5))
    assert parts[0][2] == 0 and parts[-1][3] == n
    for a, b in zip(parts, parts[1:]):
        assert a[3] == b[2]                                   # outputs: contiguous, disjoint
    speed = rate / 44100.0
    xoff = int(18 * max(1.0, speed) + 10)                     # resample.c:135-137
    for sb, se, ob, oe in parts:
        assert 0 <= sb <= se <= nframes and sb % 4 == 0
        if oe > ob and rate != 44100:
            # every output sample o reads source frames o * speed +- xoff
            assert sb <= max(0, math.floor(ob * speed) - xoff) + 4
            assert se >= min(nframes, math.ceil((oe - 1) * speed) + xoff)
        if rate == 44100:
            assert (sb, se) == (ob, oe)


def test_sums_merge_is_max_sum_min():
    a, b = longfile.new_sums(), longfile.new_sums()
    a.maxabs, a.sumsq, a.first, a.last = 3.0, 1.5, 100, 900
    b.maxabs, b.sumsq, b.first, b.last = 7.0, 2.5, 40, 500
    b.eff_first[1], b.eff_last[1] = 5, 6
    g = longfile.merge_sums([a, b])
    assert (g.maxabs, g.sumsq, g.first, g.last) == (7.0, 4.0, 40, 900)
    assert g.eff_first[1] == 5 and g.eff_last[1] == 6 and g.eff_first[0] == longfile.INT64_MAX and g.eff_last[0] == -1
    r = longfile.sums_from_array(longfile.sums_to_array(g))
    assert bytes(r) == bytes(g)


# ---- numpy part worker -----------------------------------------------------------------------------------
def _db(v):
    return math.exp(v * (math.log(10.0) / 20.0))


class NumpyJob:
    def __init__(self, an, whole, part, pcm_slice):
        self.sb, self.se, self.ob, self.oe = part
        self.x = np.asarray(pcm_slice, dtype=np.float32).reshape(-1)        # mono, analysis rate: src == out
        self.n = whole.nframes

    def peak(self):
        s = longfile.new_sums()
        s.maxabs = float(np.max(np.abs(self.x))) if self.x.size else 0.0
        s.sumsq = self.own = float(np.sum((self.x / np.float32(32768.0)).astype(np.float64) ** 2))
        return s

    def trim(self, g):
        s = api.AfxPartSums.from_buffer_copy(bytes(g))
        s.sumsq = self.own                                       # every phase's outputs merge back to the global sums
        amp = 32768.0 / g.maxabs if g.maxabs > np.float32(1e-12) else 1.0
        hit = np.nonzero(np.abs(amp * self.x.astype(np.float64)) > 32768.0 * _db(-48.0))[0]
        s.first, s.last = (int(hit[0]) + self.ob, int(hit[-1]) + self.ob) if hit.size else (longfile.INT64_MAX, -1)
        return s

    def effective(self, g):
        s = api.AfxPartSums.from_buffer_copy(bytes(g))
        s.sumsq = self.own
        lead, audible, start_off, _ = layout_of(g, self.n)
        amp = 32768.0 / g.maxabs if g.maxabs > np.float32(1e-12) else 1.0
        i = np.arange(self.ob, self.oe)
        v = np.abs(self.x.astype(np.float64) * (amp / 32768.0))
        inside = (i >= lead) & (i < lead + audible)
        for k, dbv in enumerate((-48.0, -24.0, -12.0)):
            hit = i[inside & (v > _db(dbv))]
            s.eff_first[k], s.eff_last[k] = (int(hit[0]) - lead + start_off, int(hit[-1]) - lead + start_off) if hit.size \
                else (longfile.INT64_MAX, -1)
        return s

    def read(self, begin, count, dst):
        lo, hi = max(begin, self.ob), min(begin + count, self.oe)
        if hi > lo:
            dst[lo - begin:hi - begin] = self.x[lo - self.ob:hi - self.ob]
        return max(0, hi - lo)

    def close(self):
        pass


def layout_of(g, n, N=2048):
    lead = min(g.first, n)
    trail = (n - 1 - max(g.last, lead)) if lead < n else 0
    audible = n - lead - trail
    end_off = N // 2 if (audible % N) < N // 2 else 0
    start_off = N - audible - end_off if audible + end_off < N else 0
    return lead, audible, start_off, end_off


def np_window(an, whole, g):
    lead, audible, _, _ = layout_of(g, whole.nframes)
    return lead, min(audible, 882000)


def np_finish(an, whole, g, win, begin):
    return g, win, begin


def long_mono():
    clip = synth.one_shot(77, 6.0)                     # int16 mono 44.1 kHz with leading / trailing silence
    pad = np.zeros(30000, dtype=np.int16)
    return np.concatenate([pad, clip, clip[::-1], pad, pad]).astype(np.int16)


def check_against_oracle(oracle, pcm, g, win, begin):
    data, off, pk, rms = oracle.condition(pcm)
    n = pcm.shape[0]
    lead, audible, start_off, end_off = layout_of(g, n)
    assert begin == lead and off == -lead + start_off
    assert len(data) == audible + start_off + end_off
    amp = 32768.0 / g.maxabs
    want = data[start_off:start_off + len(win)]
    assert np.array_equal(win.astype(np.float64) * (amp / 32768.0), want)
    assert abs(pk - min(1.0, g.maxabs / 32768.0)) < 1e-7
    assert abs(rms - min(1.0, math.sqrt(g.sumsq / n))) < 1e-6


def test_in_process_parts_match_oracle(oracle_lib):
    pcm = long_mono()
    whole = longfile.describe_whole(pcm.shape[0], 1, 44100, np.int16)
    for n_parts in (1, 2, 5):
        parts = longfile.plan_parts(pcm.shape[0], 44100, n_parts)
        jobs = [NumpyJob(None, whole, p, pcm[p[0]:p[1]]) for p in parts]
        g = longfile.merge_sums([j.peak() for j in jobs])
        g = longfile.merge_sums([j.trim(g) for j in jobs])
        g = longfile.merge_sums([j.effective(g) for j in jobs])
        begin, count = np_window(None, whole, g)
        win = np.zeros(count, dtype=np.float32)
        assert sum(j.read(begin, count, win) for j in jobs) == count
        check_against_oracle(oracle_lib, pcm, g, win, begin)
        whole_job = NumpyJob(None, whole, (0, pcm.shape[0], 0, pcm.shape[0]), pcm)
        ge = whole_job.effective(whole_job.trim(whole_job.peak()))
        assert list(g.eff_first) == list(ge.eff_first) and list(g.eff_last) == list(ge.eff_last)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcm = long_mono()
    whole = longfile.describe_whole(pcm.shape[0], 1, 44100, np.int16)
    part = longfile.plan_parts(pcm.shape[0], 44100, world)[rank]
    r = longfile.analyze_sharded(None, whole, part, pcm[part[0]:part[1]], dist, analysis_rank=1, job_factory=NumpyJob,
                                 finish=np_finish, window=np_window)
    dist.barrier()
    if rank == 1:
        g, win, begin = r
        torch.save({"g": longfile.sums_to_array(g), "win": win, "begin": begin}, out)
    else:
        assert r is None
    dist.destroy_process_group()


def test_two_ranks_gloo_protocol(tmp_path, oracle_lib):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    g = longfile.sums_from_array(r["g"])
    check_against_oracle(oracle_lib, long_mono(), g, r["win"], r["begin"])


@pytest.mark.parametrize("nframes,rate", [(96000 * 600, 96000), (48000 * 60, 48000), (22050 * 100, 22050), (88200 * 60, 88200),
                                          (32000 * 50, 32000), (8000 * 30, 8000), (11025 * 30, 11025), (192000 * 30, 192000),
                                          (44101 * 30, 44101), (96000 * 10 + 17, 96000), (37, 96000), (100000, 47999)])
def test_closed_form_time_replay_is_bit_exact(nframes, rate):
    """The resampler plan's O(blocks) replay of libresample's `t += dt` accumulator (resamplesubs.c:97-119) equals the
    step-by-step replay bit for bit: blocks, spans and every time checkpoint."""
    assert api.load_library().afx_debug_rs_plan_check(44100, nframes, rate) == 0
