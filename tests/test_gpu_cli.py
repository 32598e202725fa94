"""examples/main/whisper_b200_main — the Linux counterpart of the reference's CLI (Examples/main/main.cpp:174-353, textWriter.cpp): WAV in,
txt / srt / vtt out.  The transcript it writes must be the one iContext::runFull returns for the same samples (the WAV stores 16-bit
PCM, so the comparison runs the library on the de-quantised samples), in the reference's file formats (UTF-8 BOM, CRLF, SubRip
numbering and comma milliseconds, WEBVTT header)."""
import os
import subprocess
import wave

import numpy as np
import pytest

from tests.golden.make_golden import FULL_MODEL, full_pcm
from tests.test_gpu_com import open_session, run_full
from whisper_b200 import synth

pytestmark = pytest.mark.gpu
EXE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "main", "whisper_b200_main")


def fmt(t10ms, comma=False):
    ms = t10ms * 10
    return "%02d:%02d:%02d%s%03d" % (ms // 3600000, ms // 60000 % 60, ms // 1000 % 60, "," if comma else ".", ms % 1000)


def test_cli_writes_the_library_transcript(tmp_path):
    if not os.path.exists(EXE):
        pytest.fail("examples/main/whisper_b200_main is missing: run __graft_entry__.build()")
    pcm16 = np.clip(np.round(full_pcm(10)[:16000 * 50] * 32768.0), -32768, 32767).astype("<i2")
    wav = str(tmp_path / "clip.wav")
    with wave.open(wav, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm16.tobytes())
    r = subprocess.run([EXE, "-m", synth.model_path(FULL_MODEL), "-f", wav, "-otxt", "-osrt", "-ovtt", "-d", "32000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    L, h = open_session(FULL_MODEL)
    hr, segs = run_full(L, h, pcm16.astype(np.float32) / 32768.0, flags=2 | 0x40, dur=32000)
    assert hr == 0 and len(segs) >= 8
    # console: one line per segment with timestamps
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("[")]
    assert lines == ["[%s --> %s]  %s" % (fmt(s["t0"]), fmt(s["t1"]), s["text"]) for s in segs]
    txt = open(str(tmp_path / "clip.txt"), "rb").read()
    assert txt.startswith(b"\xef\xbb\xbf")
    assert txt[3:].decode().split("\r\n")[:-1] == ["[%s --> %s]  %s" % (fmt(s["t0"]), fmt(s["t1"]), s["text"].lstrip(" \t")) for s in segs]
    srt = open(str(tmp_path / "clip.srt"), "rb").read()[3:].decode()
    assert srt == "".join("%d\r\n%s --> %s\r\n%s\r\n\r\n" % (i + 1, fmt(s["t0"], True), fmt(s["t1"], True), s["text"].lstrip(" \t")) for i, s in enumerate(segs))
    vtt = open(str(tmp_path / "clip.vtt"), "rb").read()[3:].decode()
    assert vtt == "WEBVTT\r\n\r\n" + "".join("%s --> %s\r\n%s\r\n\r\n" % (fmt(s["t0"]), fmt(s["t1"]), s["text"].lstrip(" \t")) for s in segs)


def test_cli_streamed_equals_buffered(tmp_path):
    """-st: the file is read while it is transcribed (iContext::runStreamed over the CLI's block-wise WAV reader) — same console transcript
    as the buffered run of the same file (per-window mel normalisation cannot flip a decision on this clip, see test_gpu_com.py)."""
    pcm16 = np.clip(np.round(full_pcm(10)[:16000 * 50] * 32768.0), -32768, 32767).astype("<i2")
    wav = str(tmp_path / "clip.wav")
    with wave.open(wav, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm16.tobytes())
    outs = []
    for extra in ([], ["-st"]):
        r = subprocess.run([EXE, "-m", synth.model_path(FULL_MODEL), "-f", wav, "-d", "40000"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([ln for ln in r.stdout.splitlines() if ln.startswith("[")])
        assert "Memory Usage" in r.stderr and "Encode" in r.stderr          # timingsPrint's table (ContextImpl.misc.cpp:170-182)
    assert len(outs[0]) >= 8 and outs[0] == outs[1]


def test_cli_diarize_labels_the_louder_channel(tmp_path):
    """-di: a stereo file keeps its channels (createAudioBufferStereo) and every printed segment carries iContext::detectSpeaker's answer in
    the reference CLI's wording (Examples/main/main.cpp:95-117, 136); the transcript itself is that of the mono mix."""
    mono = full_pcm(10)[:16000 * 40]
    t = np.arange(mono.size) / 16000.0
    g_left = np.where((t // 4).astype(np.int64) % 2 == 0, 1.6, 0.4)          # left louder for 4 s, right louder for 4 s, ...
    st = np.empty(2 * mono.size, np.float64)
    st[0::2] = mono * g_left
    st[1::2] = mono * (2.0 - g_left)
    wav = str(tmp_path / "stereo.wav")
    with wave.open(wav, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(np.clip(np.round(st * 16384.0), -32768, 32767).astype("<i2").tobytes())
    r = subprocess.run([EXE, "-m", synth.model_path(FULL_MODEL), "-f", wav, "-d", "32000", "-di"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("[")]
    assert len(lines) >= 8 and all("(speaker " in ln for ln in lines)
    # segments inside the first 4 s are left-dominated, those inside the next 4 s right-dominated
    assert "(speaker 0)" in lines[0] and any("(speaker 1)" in ln for ln in lines)


def test_cli_rejects_bad_input(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("CLI not built")
    bad = str(tmp_path / "x.wav")
    open(bad, "wb").write(b"not a wave file")
    r = subprocess.run([EXE, "-m", synth.model_path(FULL_MODEL), bad], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 8 and "RIFF" in r.stderr
    r = subprocess.run([EXE, "-m", synth.model_path(FULL_MODEL), "-l", "klingon", bad], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 3
