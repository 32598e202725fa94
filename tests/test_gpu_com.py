"""The COM-style drop-in surface (include/whisper_b200_com.h) on a B200: loadModel -> createContext -> fullDefaultParams -> runFull ->
getResults, exactly the call sequence of the reference's CLI (Examples/main/main.cpp:210-318), driven through the flat wspc_*
helpers (ctypes cannot call C++ vtables).  The transcription driver is compared with the reference's whisper_full()
(Whisper/source/whisper.cpp:2765-3125) via fixtures generated from oracle/_ref (tests/golden/full_micro_en_ts.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.golden.make_golden import FULL_MODEL, FULL_RUNS, full_pcm
from whisper_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _lib():
    L = capi.lib()
    vp, i32 = C.c_void_p, C.c_int32
    L.wspc_open.restype = i32; L.wspc_open.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
    L.wspc_close.argtypes = [vp]
    L.wspc_run_full.restype = i32
    L.wspc_run_full.argtypes = [vp, C.POINTER(C.c_float), i32, C.c_uint32, C.c_char_p, i32, i32, i32, i32, C.POINTER(i32), i32]
    for f in ("wspc_n_segments", "wspc_n_segment_callbacks", "wspc_is_multilingual", "wspc_query_interfaces"):
        getattr(L, f).restype = i32; getattr(L, f).argtypes = [vp]
    for f in ("wspc_segment_t0", "wspc_segment_t1"):
        getattr(L, f).restype = C.c_int64; getattr(L, f).argtypes = [vp, i32]
    L.wspc_segment_text.restype = C.c_char_p; L.wspc_segment_text.argtypes = [vp, i32]
    L.wspc_segment_n_tokens.restype = i32; L.wspc_segment_n_tokens.argtypes = [vp, i32]
    L.wspc_token_id.restype = i32; L.wspc_token_id.argtypes = [vp, i32, i32]
    L.wspc_token_p.restype = C.c_float; L.wspc_token_p.argtypes = [vp, i32, i32]
    L.wspc_token_flags.restype = i32; L.wspc_token_flags.argtypes = [vp, i32, i32]
    L.wspc_tokenize.restype = i32; L.wspc_tokenize.argtypes = [vp, C.c_char_p, C.POINTER(i32), i32]
    L.wspc_string_from_token.restype = C.c_char_p; L.wspc_string_from_token.argtypes = [vp, i32]
    L.wspc_special_tokens.restype = i32; L.wspc_special_tokens.argtypes = [vp, C.POINTER(i32)]
    L.wspc_find_language_key.restype = C.c_uint32; L.wspc_find_language_key.argtypes = [C.c_char_p]
    L.wspc_language_count.restype = i32
    return L


@pytest.fixture(scope="module")
def session():
    L = _lib()
    h = C.c_void_p()
    hr = L.wspc_open(synth.model_path(FULL_MODEL).encode(), 0, C.byref(h))
    assert hr == 0, hex(hr & 0xFFFFFFFF)
    yield L, h
    L.wspc_close(h)


def run_full(L, h, pcm, flags=0, language=b"en", max_tokens=0, threads=4, off=0, dur=0, prompt=None):
    pcm = np.ascontiguousarray(pcm, np.float32)
    pt = None if prompt is None else np.ascontiguousarray(prompt, np.int32)
    hr = L.wspc_run_full(h, pcm.ctypes.data_as(C.POINTER(C.c_float)), pcm.size, flags, language, max_tokens, threads, off, dur,
                         None if pt is None else pt.ctypes.data_as(C.POINTER(C.c_int32)), 0 if pt is None else pt.size)
    segs = []
    for i in range(L.wspc_n_segments(h)):
        n = L.wspc_segment_n_tokens(h, i)
        segs.append(dict(t0=L.wspc_segment_t0(h, i), t1=L.wspc_segment_t1(h, i), text=L.wspc_segment_text(h, i).decode(errors="replace"),
                         tokens=[L.wspc_token_id(h, i, j) for j in range(n)], flags=[L.wspc_token_flags(h, i, j) for j in range(n)]))
    return hr, segs


@pytest.mark.parametrize("name", list(FULL_RUNS))
def test_run_full_matches_reference_driver(session, name):
    L, h = session
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "full_micro_en_ts.npz"))
    flags, max_tokens, off, dur = FULL_RUNS[name]
    # eFullParamsFlags::NoContext is implied between parametrised runs: the reference fixture was made with a fresh context each time
    hr, segs = run_full(L, h, full_pcm(), flags=flags | 2, max_tokens=max_tokens, off=off, dur=dur)
    assert hr == 0
    ref_t = g[name + "_t"]
    assert len(segs) == len(ref_t)
    assert [[s["t0"], s["t1"]] for s in segs] == ref_t.tolist()
    assert [len(s["tokens"]) for s in segs] == g[name + "_ntok"].tolist()
    assert [t for s in segs for t in s["tokens"]] == g[name + "_tokens"].tolist()
    assert [s["text"] for s in segs] == g[name + "_text"].tolist()
    assert L.wspc_n_segment_callbacks(h) == len(segs)     # new_segment_callback fired once per segment
    for s in segs:                                        # eTokenFlags::Special <=> id >= eot (convertThings.cpp:200-203)
        assert s["flags"] == [1 if t >= 50256 else 0 for t in s["tokens"]]


def test_short_audio_is_a_noop(session):
    L, h = session
    hr, segs = run_full(L, h, synth.synth_pcm(0, 8000))      # < 1 s: whisper_full returns without doing anything (whisper.cpp:2813-2818)
    assert hr == 0 and segs == []
    hr, segs = run_full(L, h, np.zeros(0, np.float32))
    assert hr == 0 and segs == []


def test_unknown_language_and_unsupported_flags(session):
    L, h = session
    hr, _ = run_full(L, h, synth.synth_pcm(0, 32000), language=b"zz")
    assert (hr & 0xFFFFFFFF) == 0x80070057                     # E_INVALIDARG
    hr, _ = run_full(L, h, synth.synth_pcm(0, 32000), flags=0x200)   # SpeedupAudio
    assert (hr & 0xFFFFFFFF) == 0x80004001                     # E_NOTIMPL


def test_imodel_surface(session):
    L, h = session
    assert L.wspc_query_interfaces(h) == 0
    assert L.wspc_is_multilingual(h) == 0
    st = (C.c_int32 * 8)()
    assert L.wspc_special_tokens(h, st) == 0
    assert list(st) == [50256, 50257, 50360, 50361, 50362, 50363, 50358, 50359]
    assert L.wspc_string_from_token(h, 42) == b" t42"
    assert L.wspc_string_from_token(h, 50363) == b"[_BEG_]"
    # the synthetic vocabulary is " t<i>": the GPT-2 pre-split separates letters from digits, so nothing matches (as in the reference)
    out = (C.c_int32 * 16)()
    assert L.wspc_tokenize(h, b" t12 t345", out, 16) == 0
    assert L.wspc_language_count() == 99
    assert L.wspc_find_language_key(b"en") == ord("e") | (ord("n") << 8)
    assert L.wspc_find_language_key(b"haw") == ord("h") | (ord("a") << 8) | (ord("w") << 16)
    assert L.wspc_find_language_key(b"english") == ord("e") | (ord("n") << 8)
    assert L.wspc_find_language_key(b"klingon") == 0xFFFFFFFF


# iModel::tokenize against the reference's whisper_tokenize (whisper.cpp:2192-2245, 2378-2391) on the "-words" vocabulary
# (synth.tokenizer_test_words).  Expected ids were produced by oracle/_ref (RefOracle.tokenize) with tests/golden/make_golden.py
# --tokenizer; they include the reference's quirk of emitting one single-character token right after every non-final match.
TOKENIZER_GOLDEN = {
    b" hello hellox 12345!": [100, 100, 23, 103, 29, 30, 31],
    b" hello world's , worlds 345 12": [100, 109, 107, 108, 109, 18, 36, 29, 30, 31, 103],
    b" t12 t345": [36, 19, 27, 28, 36, 19, 104],
    b"": [],
}


def test_tokenize_matches_reference():
    L = _lib()
    h = C.c_void_p()
    assert L.wspc_open(synth.model_path("micro.en-words").encode(), 0, C.byref(h)) == 0
    try:
        out = (C.c_int32 * 64)()
        for text, want in TOKENIZER_GOLDEN.items():
            n = L.wspc_tokenize(h, text, out, 64)
            assert n == len(want) and list(out)[:n] == want, (text, list(out)[:max(n, 0)])
    finally:
        L.wspc_close(h)
