mkdir -p gpurun_out
(timeout 800 python -m pytest tests -m gpu -q -x 2>&1 | tail -3)
python - <<EOF
import sys; sys.path.insert(0,".")
from afec_b200 import api, synth
pcms = synth.tiled_corpus(400, 16, seconds=3.0, seed0=0)
an = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_ALL)
b = an.batch(pcms, [44100]*len(pcms)); b.upload()
for _ in range(3): b.compute()
b.sync()
print("multi-stream compute ms:", b.timings()[1])
EOF
AFX_SINGLE_STREAM=1 python - <<EOF
import sys; sys.path.insert(0,".")
from afec_b200 import api, synth
pcms = synth.tiled_corpus(400, 16, seconds=3.0, seed0=0)
an = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_ALL)
b = an.batch(pcms, [44100]*len(pcms)); b.upload()
for _ in range(3): b.compute()
b.sync()
print("single-stream compute ms:", b.timings()[1])
EOF
