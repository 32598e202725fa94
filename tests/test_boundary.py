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


def test_cli_streaming_wav_reader(tmp_path):
    """The CLI's block-wise WAV reader (the pull source it hands to iContext::runStreamed) delivers exactly the samples of its buffered
    reader: 16-bit mono, 16-bit stereo (down-mixed), 32-bit float with a foreign chunk of odd length in front of the data.  No GPU: the
    CLI's --verify-stream self-check loads no model."""
    import struct
    import wave
    exe = os.path.join(os.path.dirname(HERE), "examples", "main", "whisper_b200_main")
    if not os.path.exists(exe):
        pytest.fail("examples/main/whisper_b200_main is missing: run __graft_entry__.build()")
    rng = np.random.default_rng(5)
    paths = []
    for name, ch, n in (("mono.wav", 1, 16000 * 7 + 13), ("stereo.wav", 2, 16000 * 3 + 1)):
        p = str(tmp_path / name)
        with wave.open(p, "wb") as w:
            w.setnchannels(ch); w.setsampwidth(2); w.setframerate(16000)
            w.writeframes(rng.integers(-30000, 30000, size=n * ch).astype("<i2").tobytes())
        paths.append(p)
    data = rng.standard_normal(50001).astype("<f4").tobytes()
    body = (b"WAVE" + b"fmt " + struct.pack("<I", 16) + struct.pack("<HHIIHH", 3, 1, 16000, 64000, 4, 32)
            + b"LIST" + struct.pack("<I", 5) + b"abcde\0" + b"data" + struct.pack("<I", len(data)) + data)
    p = str(tmp_path / "float.wav")
    open(p, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    paths.append(p)
    out = _run(exe, "--verify-stream", *sum((["-f", q] for q in paths), []))
    assert out.count("streamed reader identical") == 3


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
