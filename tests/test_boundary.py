"""The drop-in boundary, proven against the reference's OWN headers (Whisper/API/*.h + ComLightLib/):

* layout: tests/boundary/layout_probe.cpp prints ~110 ABI facts (struct sizes, field offsets, enum values, interface GUID bytes, vtable
  slot numbers).  Compiled against include/whisper_b200_com.h it must print exactly what it prints when compiled against the reference
  headers — live when /root/reference is present, else against the committed copy of that output (tests/boundary/layout_reference.txt).
* behaviour (GPU): tests/boundary/ref_client.cpp is a client application compiled ONLY against the reference headers (it includes nothing
  from this repository), linked with libwhisper_b200.so.  It makes the call sequence of the reference's CLI (Examples/main/main.cpp:210-318)
  with its own iAudioBuffer object; its transcript must equal the reference's whisper_full fixture (tests/golden/full_runs.npz).
The binaries are built by tests/boundary/Makefile (from __graft_entry__.build()) and travel to the GPU box."""
import os
import subprocess

import numpy as np
import pytest

from whisper_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "boundary", "_build")


def _run(path, *args):
    r = subprocess.run([path, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    return r.stdout


def test_layout_matches_reference_headers():
    ours = os.path.join(BUILD, "probe_ours")
    if not os.path.exists(ours):
        pytest.fail("tests/boundary/_build/probe_ours is missing: run __graft_entry__.build()")
    got = _run(ours)
    want = open(os.path.join(HERE, "boundary", "layout_reference.txt")).read()
    assert len(want.splitlines()) > 100
    assert got == want
    ref = os.path.join(BUILD, "probe_ref")
    if os.path.exists(ref) and os.path.isdir("/root/reference"):
        assert _run(ref) == want, "layout_reference.txt is stale: regenerate it with tests/boundary/_build/probe_ref"


def test_pcm_streamer_unit():
    """The PCM queue behind iContext::runStreamed (whisper_b200/csrc/pcm_streamer.h, host-only): windows at increasing offsets are the
    source's samples for any block sizes, with and without the background reader; backward seeks and source failures are reported."""
    exe = os.path.join(BUILD, "streamer_test")
    if not os.path.exists(exe):
        pytest.fail("tests/boundary/_build/streamer_test is missing: run __graft_entry__.build()")
    assert "streamer_test: ok" in _run(exe)


def test_capture_loop_unit():
    """The listening loop behind iContext::runCapture (whisper_b200/csrc/capture_loop.h + vad.h, host-only) with a fake transcriber:
    utterance cutting by the reference's rules, intact samples at the claimed offsets, status reports, stall-and-drop under a slow
    transcriber, error propagation, E_EOF at the end of the source."""
    exe = os.path.join(BUILD, "capture_test")
    if not os.path.exists(exe):
        pytest.fail("tests/boundary/_build/capture_test is missing: run __graft_entry__.build()")
    assert "capture_test: ok" in _run(exe)


def test_capture_example_builds_and_parses_its_options():
    """examples/capture/whisper_b200_capture (live transcription from a pipe over createAudioCapture + runCapture): built, prints its usage,
    rejects an unknown language before touching the GPU.  (Its transcription path is iContext::runCapture, tested in test_gpu_com.py.)"""
    exe = os.path.join(os.path.dirname(HERE), "examples", "capture", "whisper_b200_capture")
    if not os.path.exists(exe):
        pytest.fail("examples/capture/whisper_b200_capture is missing: run __graft_entry__.build()")
    r = subprocess.run([exe, "-h"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0 and "s16le 16 kHz mono" in r.stderr
    r = subprocess.run([exe, "-l", "klingon"], stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 3 and "unknown language" in r.stderr


def _wav_cases(tmp_path):
    import struct
    import wave
    rng = np.random.default_rng(11)
    cases = []
    for name, ch, rate, n in (("mono16k.wav", 1, 16000, 16000 * 3 + 7), ("stereo44k.wav", 2, 44100, 44100 * 2 + 3), ("mono8k.wav", 1, 8000, 8000 * 2 + 1),
                              ("four48k.wav", 4, 48000, 48000 + 5)):
        x = rng.integers(-30000, 30000, size=(n, ch)).astype("<i2")
        p = str(tmp_path / name)
        with wave.open(p, "wb") as w:
            w.setnchannels(ch); w.setsampwidth(2); w.setframerate(rate); w.writeframes(x.tobytes())
        cases.append((p, x.astype(np.float32) / np.float32(32768.0), rate))
    f = rng.standard_normal((22050 + 9, 2)).astype("<f4")
    body = (b"WAVE" + b"fmt " + struct.pack("<I", 16) + struct.pack("<HHIIHH", 3, 2, 22050, 22050 * 8, 8, 32)
            + b"LIST" + struct.pack("<I", 3) + b"abc\0" + b"data" + struct.pack("<I", f.nbytes) + f.tobytes())
    p = str(tmp_path / "float22k.wav")
    open(p, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    cases.append((p, f.astype(np.float32), 22050))
    return cases


def _expected_16k(x, rate):
    """(mono, first two channels) at 16 kHz: channel average, then — for other rates — the linear interpolation wav_reader.h documents."""
    mono = (x.sum(axis=1, dtype=np.float32) if x.shape[1] <= 2 else np.add.reduce(x, axis=1, dtype=np.float32)) / np.float32(x.shape[1])
    if x.shape[1] > 2:          # the decoder adds channel by channel in f32
        acc = np.zeros(x.shape[0], np.float32)
        for c in range(x.shape[1]):
            acc = acc + x[:, c]
        mono = acc / np.float32(x.shape[1])
    lr = x[:, :2] if x.shape[1] >= 2 else np.repeat(mono[:, None], 2, axis=1)
    if rate == 16000:
        return mono, lr
    n = x.shape[0]
    m = int(float(n) * 16000.0 / rate)
    pos = np.arange(m, dtype=np.float64) * rate / 16000.0
    i0 = pos.astype(np.int64)
    i1 = np.minimum(i0 + 1, n - 1)
    t = (pos - i0).astype(np.float32)
    one = np.float32(1.0)
    return mono[i0] * (one - t) + mono[i1] * t, lr[i0] * (one - t)[:, None] + lr[i1] * t[:, None]


def test_media_layer_decodes_wav_files(tmp_path):
    """initMediaFoundation -> loadAudioFile / openAudioFile / loadAudioFileData (the calls of the reference CLI, Examples/main/main.cpp:306-318)
    on RIFF/WAVE input: 16-bit and float, 1 / 2 / 4 channels, 8 - 48 kHz -> 16 kHz mono (+ stereo pairs), the same samples whether the file
    is loaded whole, streamed from disk or streamed from memory in blocks of any size; announced duration = delivered samples.  No GPU."""
    import ctypes as C
    from whisper_b200 import capi
    capi.lib()
    shim = os.path.join(BUILD, "libwspc_test.so")
    if not os.path.exists(shim):
        pytest.fail("tests/boundary/_build/libwspc_test.so is missing: run __graft_entry__.build()")
    L = C.CDLL(shim)
    fp = C.POINTER(C.c_float)
    L.wspc_media_decode.restype = C.c_int32
    L.wspc_media_decode.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, fp, fp, C.c_int32, C.POINTER(C.c_int32)]
    for path, x, rate in _wav_cases(tmp_path):
        want_mono, want_lr = _expected_16k(x, rate)
        cap = want_mono.size + 1000
        mono, stereo, info = np.zeros(cap, np.float32), np.zeros(2 * cap, np.float32), (C.c_int32 * 3)()
        n = L.wspc_media_decode(path.encode(), 0, 1, 0, mono.ctypes.data_as(fp), stereo.ctypes.data_as(fp), cap, info)
        assert n == want_mono.size, (path, n)
        assert np.abs(mono[:n] - want_mono).max() < 1e-6
        assert info[0] == (1 if x.shape[1] >= 2 else 0)
        if x.shape[1] >= 2:
            assert np.abs(stereo[:2 * n].reshape(n, 2) - want_lr).max() < 1e-6
        # without the stereo request no pairs are kept
        n0 = L.wspc_media_decode(path.encode(), 0, 0, 0, mono.ctypes.data_as(fp), stereo.ctypes.data_as(fp), cap, info)
        assert n0 == n and info[0] == 0
        for mode, block in ((1, 777), (1, 100000), (2, 4096)):
            got = np.zeros(cap, np.float32)
            k = L.wspc_media_decode(path.encode(), mode, 1, block, got.ctypes.data_as(fp), None, cap, info)
            assert k == n and np.array_equal(got[:n], mono[:n]), (path, mode, block)
            ticks = (info[1] & 0xFFFFFFFF) | (info[2] << 32)
            assert ticks == n * 625 and info[0] == (1 if x.shape[1] >= 2 else 0)
    # failures: a missing file, a file that is not RIFF/WAVE, an unsupported sample format
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"definitely not audio")
    buf = np.zeros(16, np.float32)
    assert (L.wspc_media_decode(str(tmp_path / "missing.wav").encode(), 0, 0, 0, buf.ctypes.data_as(fp), None, 16, None) & 0xFFFFFFFF) == 0x80070002
    assert (L.wspc_media_decode(str(bad).encode(), 1, 0, 64, buf.ctypes.data_as(fp), None, 16, None) & 0xFFFFFFFF) == 0x80070057
    import struct
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + struct.pack("<HHIIHH", 1, 1, 16000, 16000, 1, 8) + b"data" + struct.pack("<I", 4) + b"\0\0\0\0"
    (tmp_path / "u8.wav").write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert (L.wspc_media_decode(str(tmp_path / "u8.wav").encode(), 0, 0, 0, buf.ctypes.data_as(fp), None, 16, None) & 0xFFFFFFFF) == 0x80070057


def test_com_exports_include_the_streaming_factory():
    """whisper.def's exports plus the three Linux factories (createAudioBuffer, createAudioReader, createAudioCapture) are in the product library."""
    from whisper_b200 import capi
    so = capi.lib()._name
    syms = subprocess.run(["nm", "-DC", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    for f in ("setupLogger", "loadModel", "findLanguageKeyW", "findLanguageKeyA", "getSupportedLanguages", "listGPUs", "initMediaFoundation",
              "createAudioBuffer", "createAudioReader", "createAudioCapture"):
        assert "Whisper::%s(" % f in syms, f


def _client_segments(out):
    segs = []
    for line in out.splitlines():
        if line.startswith("seg "):
            head, rest = line.split(" [", 1)
            toks, text = rest.split("] ", 1) if "] " in rest else (rest.rstrip("]"), "")
            _, t0, t1 = head.split()
            segs.append((int(t0), int(t1), [int(x) for x in toks.split()] if toks.strip() else [], text))
    return segs


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plain", "context_second_call", "ml_translate", "plain/stream"])
def test_client_built_against_reference_headers(name, tmp_path):
    from tests.golden.make_golden import FULL_RUNS, full_pcm
    exe = os.path.join(BUILD, "ref_client")
    if not os.path.exists(exe):
        pytest.skip("ref_client was not built (needs the reference headers at build time)")
    g = np.load(os.path.join(HERE, "golden", "full_runs.npz"))
    name, _, mode = name.partition("/")          # "/stream": the same clip through iContext::runStreamed from the client's own iAudioReader
    model, flags, max_tokens, off, dur, lang, calls = FULL_RUNS[name]
    assert max_tokens == 0 and off == 0
    pcm_path = str(tmp_path / "clip.f32")
    full_pcm(int(g[name + "_pcm_base"])).astype("<f4").tofile(pcm_path)
    out = _run(exe, synth.model_path(model), pcm_path, str(flags), lang, str(calls), str(dur), *([mode] if mode else []))
    segs = _client_segments(out)
    assert [[s[0], s[1]] for s in segs] == g[name + "_t"].tolist()
    assert [len(s[2]) for s in segs] == g[name + "_ntok"].tolist()
    assert [t for s in segs for t in s[2]] == g[name + "_tokens"].tolist()
    assert [s[3] for s in segs] == g[name + "_text"].tolist()
    assert "callbacks %d" % (len(segs) if calls == 1 else 0) in out or calls > 1
