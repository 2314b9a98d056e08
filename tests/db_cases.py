"""The small file set behind tests/golden/ref_ll.db (reference-written afec-ll.db)."""
import os

import numpy as np

from afec_b200 import synth
from oracle import oracle


def cases():
    """name -> (int16 pcm, rate) or raw bytes for a broken file."""
    return {
        "kick.wav": (synth.one_shot(1, 0.6), 44100),
        "pad_stereo.wav": (synth.one_shot(4, 0.5, channels=2), 44100),
        "hat_48k.wav": (synth.one_shot(21, 0.35, rate=48000), 48000),
        "_Not A Wavefile.wav": b"this is not a wave file",
    }


def write_files(directory: str) -> list:
    paths = []
    for name, c in cases().items():
        p = os.path.join(directory, name)
        if isinstance(c, bytes):
            with open(p, "wb") as f:
                f.write(c)
        else:
            oracle.write_wav(p, c[0], c[1])
        paths.append(p)
    return paths
