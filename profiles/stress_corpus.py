"""Edge-material corpus (tests / sweeps only): the signals a sample library really holds next to decaying one-shots -- steady
tones on and off bin centres, square / saw / impulse trains, DC, full-scale and clipped noise, Nyquist tones, chirps, one
impulse in silence, silence next to bursts, phase-inverted stereo (a mono sum of zero), few-LSB material.

    python profiles/stress_corpus.py self [hop]     CPU: the oracle with its two FFT variants, under tests/parity.py's rules
    python profiles/stress_corpus.py gpu [hop]      CUDA path against the oracle, same rules"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


from edge_corpus import build


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "self"
    hop = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    import parity
    from oracle import oracle
    oracle.build()
    files = build()
    bad = 0
    got = None
    if mode == "gpu":
        from afec_b200 import api
        an = api.SampleAnalyser(44100, 2048, hop, features=api.FEAT_ALL)
        got = an.analyze_pcm([p for _, p, _ in files], [r for _, _, r in files])
    t0 = time.time()
    for i, (name, p, r) in enumerate(files):
        want = oracle.analyze(p, src_rate=r, hop=hop, file_size=44 + p.size * 2)
        data = oracle.condition(p, src_rate=r)[0] if want.status == 0 else None
        if mode == "gpu":
            other = got[i]
        else:
            oracle.set_fft_variant(1)
            try:
                other = oracle.analyze(p, src_rate=r, hop=hop, file_size=44 + p.size * 2)
            finally:
                oracle.set_fft_variant(0)
        errs = parity.compare(other, want, mdata=data, hop=hop)
        if errs:
            bad += 1
            print("%-20s status %d F %d: %d mismatches; first: %s" % (name, want.status, want.F, len(errs), errs[:3]), flush=True)
    print("stress corpus (%s): %d files, hop %d, %.0f s; files with mismatches: %d" % (mode, len(files), hop, time.time() - t0, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
