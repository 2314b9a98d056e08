#!/bin/bash
# Round 2, call T: FFT-form autocorrelation (k_autocorr<2>): parity tests, A/B kernel times against the direct FP32 sums,
# 320-file sweep at hop 1024 and 160 at hop 512, accuracy of the FFT form against the FP64 kernel on the sweep corpus.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2t_tests.log; cat gpurun_out/r2t_tests.log
for v in 1 0; do VT_MIXED=1 AFX_AUTOCORR_DIRECT=$v timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/r2t_variant_direct_$v.log 2>&1; tail -1 gpurun_out/r2t_variant_direct_$v.log; done
(timeout 900 python profiles/parity_sweep.py 320 1024 7000 2>&1 | tail -6) > gpurun_out/r2t_sweep_1024.log; cat gpurun_out/r2t_sweep_1024.log
(timeout 600 python profiles/parity_sweep.py 160 512 9000 2>&1 | tail -6) > gpurun_out/r2t_sweep_512.log; cat gpurun_out/r2t_sweep_512.log
cat > /tmp/acc.py <<'P'
import os, sys, subprocess, numpy as np
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    from afec_b200 import api, synth
    rng = np.random.default_rng(5)
    pcms = [synth.one_shot(3000 + i, float(np.exp(rng.uniform(np.log(0.1), np.log(20.0))))) for i in range(200)]
    pcms += [(p.astype(np.float64) * 0.002).astype(np.int16) for p in pcms[:20]] + [(p // 4 + 9000).astype(np.int16) for p in pcms[20:40]]
    an = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_AUTOCORR | api.FEAT_STATS)
    r = an.analyze_pcm(pcms, [44100] * len(pcms))
    np.save(sys.argv[1], np.concatenate([np.asarray(x.series("auto_correlation")).reshape(-1) for x in r]))
else:
    for name, env in (("fp64", {"AFX_AUTOCORR_FP64": "1"}), ("fft", {})):
        subprocess.check_call([sys.executable, "/tmp/acc.py", "/tmp/ac_%s.npy" % name], env=dict(os.environ, **env))
    a, b = np.load("/tmp/ac_fp64.npy"), np.load("/tmp/ac_fft.npy")
    err = np.abs(a - b); tol = 1e-6 + 1e-4 * np.abs(a)
    print("autocorrelation, FFT form vs FP64 kernel: %d frames, max |d| %.3g, max |d| / tol %.3g, frames over tol %d" % (a.size, err.max(), (err / tol).max(), int((err > tol).sum())))
P
timeout 600 python /tmp/acc.py > gpurun_out/r2t_accuracy.log 2>&1; tail -3 gpurun_out/r2t_accuracy.log
