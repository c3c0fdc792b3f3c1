"""Host tests of ssspy_b200.io (wavread / wavwrite): byte-exact against files written by the reference
(tests/golden/io_wav.npz, made by tests/golden/make_golden_io.py) and the cases of the reference's
tests/package/io/test_wavread.py (scipy cross-check, round trips, num_frames bound, invalid headers)."""
import os
import struct

import numpy as np
import pytest

from ssspy_b200 import wavread, wavwrite

GOLD = os.path.join(os.path.dirname(__file__), "golden", "io_wav.npz")
CASES = ["mono_f64", "mono_f32", "mono2d_f64", "stereo_f64", "stereo_cf_f64", "stereo_i16", "mono_i8"]


def _opt(v, as_bool=False):
    return None if v < 0 else (bool(v) if as_bool else int(v))


@pytest.mark.parametrize("name", CASES)
def test_wavwrite_bytes_match_reference(name, tmp_path):
    g = np.load(GOLD)
    path = str(tmp_path / "out.wav")
    wavwrite(path, g[name + "/in"], int(g[name + "/rate"]), channels_first=_opt(int(g[name + "/cf"]), True))
    with open(path, "rb") as f:
        assert f.read() == g[name + "/bytes"].tobytes()


@pytest.mark.parametrize("name", CASES)
def test_wavread_matches_reference(name, tmp_path):
    g = np.load(GOLD)
    path = str(tmp_path / "in.wav")
    with open(path, "wb") as f:
        f.write(g[name + "/bytes"].tobytes())
    for k, (off, num, r2d, rcf) in enumerate(g["reads"]):
        data, rate = wavread(path, frame_offset=int(off), num_frames=_opt(num), return_2d=_opt(r2d, True),
                             channels_first=_opt(rcf, True))
        want = g["{}/read{}".format(name, k)]
        assert rate == int(g[name + "/rate"])
        assert data.dtype == np.float64 and data.shape == want.shape
        assert np.array_equal(data, want)


@pytest.mark.parametrize("n_channels", [1, 2])
@pytest.mark.parametrize("frame_offset", [0, 10])
@pytest.mark.parametrize("num_frames", [None, 100])
def test_wavread_against_scipy(n_channels, frame_offset, num_frames, tmp_path):
    wavfile = pytest.importorskip("scipy.io.wavfile")
    rng = np.random.default_rng(1)
    shape = (4000,) if n_channels == 1 else (4000, n_channels)
    pcm = rng.integers(-2**15, 2**15, size=shape, dtype="<i2")
    path = str(tmp_path / "scipy.wav")
    wavfile.write(path, 16000, pcm)
    data, rate = wavread(path, frame_offset=frame_offset, num_frames=num_frames)
    end = None if num_frames is None else frame_offset + num_frames
    assert rate == 16000
    assert np.array_equal(data, pcm[frame_offset:end] / 2**15)


@pytest.mark.parametrize("is_float", [True, False])
@pytest.mark.parametrize("n_channels", [0, 1, 2])
@pytest.mark.parametrize("channels_first", [True, False, None])
def test_wavio_round_trip(is_float, n_channels, channels_first, tmp_path):
    rng = np.random.default_rng(0)
    n = 8000
    shape = (n,) if n_channels == 0 else ((n_channels, n) if channels_first else (n, n_channels))
    pcm = rng.integers(-2**15, 2**15, size=shape, dtype="<i2")
    given = pcm / 2**15 if is_float else pcm
    path = str(tmp_path / "valid.wav")
    if n_channels == 0:
        wavwrite(path, given, sample_rate=16000)
        back, _ = wavread(path)
    else:
        wavwrite(path, given, sample_rate=16000, channels_first=channels_first)
        back, _ = wavread(path, return_2d=True, channels_first=channels_first)
    assert np.array_equal(back, pcm / 2**15)


@pytest.mark.parametrize("n_channels", [1, 2])
@pytest.mark.parametrize("frame_offset", [0, 10])
def test_wavread_num_frames_bound(n_channels, frame_offset, tmp_path):
    max_frame = 1000
    path = str(tmp_path / "bound.wav")
    wavwrite(path, np.zeros((max_frame, n_channels)), 16000)
    wavread(path, frame_offset=frame_offset, num_frames=max_frame - frame_offset)
    bad = max_frame - frame_offset + 1
    with pytest.raises(ValueError) as e:
        wavread(path, frame_offset=frame_offset, num_frames=bad)
    assert str(e.value) == "num_frames={} exceeds maximum frame {}.".format(bad, max_frame)
    with pytest.raises(ValueError):
        wavread(path, num_frames=-1)


def test_wavread_invalid_metadata(tmp_path):
    good = str(tmp_path / "good.wav")
    wavwrite(good, np.zeros(160), 16000)
    with open(good, "rb") as f:
        head = bytearray(f.read())

    def check(offset, patch, exc, message):
        broken = bytearray(head)
        broken[offset:offset + len(patch)] = patch
        path = str(tmp_path / "broken.wav")
        with open(path, "wb") as f:
            f.write(bytes(broken))
        with pytest.raises(exc) as e:
            wavread(path)
        assert str(e.value) == message

    check(0, b"RIFX", NotImplementedError, "Not support {}.".format(b"RIFX"))
    check(8, b"wave", NotImplementedError, "Not support {}.".format(b"wave"))
    check(12, b"FMT ", NotImplementedError, "Not support {}.".format(b"FMT "))
    check(16, struct.pack("<I", 15), NotImplementedError, "Invalid header is detected.")
    check(20, struct.pack("<H", 0), NotImplementedError, "Invalid header 0 is detected.")
    check(28, struct.pack("<I", 1), ValueError, "Invalid header is detected.")
    check(36, b"DATA", NotImplementedError, "Not support {}.".format(b"DATA"))


def test_wavwrite_invalid_input(tmp_path):
    path = str(tmp_path / "x.wav")
    with pytest.raises(ValueError):
        wavwrite(path, np.zeros((10, 3)), 16000)
    with pytest.raises(ValueError):
        wavwrite(path, np.zeros((2, 2, 2)), 16000)
    with pytest.raises(ValueError):
        wavwrite(path, np.zeros(10, dtype=np.int32), 16000)
    with pytest.raises(AssertionError):
        wavwrite(str(tmp_path / "x.flac"), np.zeros(10), 16000)
