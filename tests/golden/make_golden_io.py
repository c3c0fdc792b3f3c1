"""Generates tests/golden/io_wav.npz by running the REFERENCE's ssspy.io.wavwrite / wavread (ssspy/io/__init__.py)
in this container.  Each case stores the waveform given to wavwrite, the exact file bytes the reference wrote and what
the reference reads back for a set of (frame_offset, num_frames, return_2d, channels_first) argument tuples.

Run from the repo root:  python tests/golden/make_golden_io.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, "/root/reference")
from ssspy.io import wavread, wavwrite  # noqa: E402

READS = [(0, None, None, None), (10, 50, None, None), (0, None, True, None), (0, None, True, True), (7, None, None, True)]


def main():
    rng = np.random.default_rng(20240607)
    cases = {
        "mono_f64": (rng.uniform(-1, 0.999, 400), None),
        "mono_f32": (rng.uniform(-1, 0.999, 333).astype(np.float32), None),
        "mono2d_f64": (rng.uniform(-1, 0.999, (256, 1)), False),
        "stereo_f64": (rng.uniform(-1, 0.999, (300, 2)), None),
        "stereo_cf_f64": (rng.uniform(-1, 0.999, (2, 300)), True),
        "stereo_i16": (rng.integers(-32768, 32768, (200, 2)).astype("<i2"), None),
        "mono_i8": (rng.integers(-128, 128, 150).astype("i1"), None),
    }
    out = {"reads": np.array([[-1 if v is None else int(v) for v in r] for r in READS])}
    with tempfile.TemporaryDirectory() as d:
        for name, (wav, cf) in cases.items():
            path = os.path.join(d, name + ".wav")
            rate = 8000 if "i8" in name else 16000
            wavwrite(path, wav, rate, channels_first=cf)
            out[name + "/in"] = wav
            out[name + "/cf"] = np.array(-1 if cf is None else int(cf))
            out[name + "/rate"] = np.array(rate)
            with open(path, "rb") as f:
                out[name + "/bytes"] = np.frombuffer(f.read(), dtype=np.uint8)
            for k, (off, num, r2d, rcf) in enumerate(READS):
                data, sr = wavread(path, frame_offset=off, num_frames=num, return_2d=r2d, channels_first=rcf)
                assert sr == rate
                out["{}/read{}".format(name, k)] = np.array(data)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "io_wav.npz"), **out)


if __name__ == "__main__":
    main()
