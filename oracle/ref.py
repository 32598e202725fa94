"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/liboracle_ref.so (the reference's own CPU implementation,
Whisper/source/{ggml.c,whisper.cpp}, compiled unmodified by oracle/Makefile).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# WSP_ORACLE_LIB selects another build of the same sources (e.g. the scalar, non-AVX build `make scalar` produces, used to measure
# the reference's own build-to-build spread; see tests/golden/make_golden.py --spread)
LIB_PATH = os.environ.get("WSP_ORACLE_LIB") or os.path.join(_HERE, "_ref", "liboracle_ref.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, ci, fp, ip = C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.ora_init.restype = vp; L.ora_init.argtypes = [C.c_char_p]
        L.ora_free.argtypes = [vp]
        L.ora_system_info.restype = C.c_char_p
        L.ora_hparams.argtypes = [vp, ip]
        L.ora_special_tokens.argtypes = [vp, ip]
        L.ora_pcm_to_mel.argtypes = [vp, fp, ci, ci]
        L.ora_set_mel.argtypes = [vp, fp, ci, ci]
        L.ora_mel_len.argtypes = [vp]
        L.ora_get_mel.argtypes = [vp, fp]
        L.ora_encode.argtypes = [vp, ci, ci]
        L.ora_decode.argtypes = [vp, ip, ci, ci, ci]
        L.ora_logits_size.argtypes = [vp]
        L.ora_get_logits.argtypes = [vp, fp]
        L.ora_get_probs.argtypes = [vp, fp]
        L.ora_sample_best.argtypes = [vp, ip, fp]
        L.ora_sample_timestamp.argtypes = [vp, ci, ip, fp]
        L.ora_sample_from_probs.argtypes = [vp, fp, ci, ci, ip, fp]
        L.ora_lang_auto_detect.restype = ci; L.ora_lang_auto_detect.argtypes = [vp, ci, ci, fp]
        L.ora_lang_max_id.restype = ci
        L.ora_gap_log_enable.argtypes = [vp, ci]
        L.ora_gap_log_get.argtypes = [fp]
        L.ora_tokenize.restype = ci; L.ora_tokenize.argtypes = [vp, C.c_char_p, ip, ci]
        L.ora_cross_kv_elements.restype = C.c_int64; L.ora_cross_kv_elements.argtypes = [vp]
        L.ora_get_cross_kv.argtypes = [vp, fp, fp]
        L.ora_self_kv_elements.restype = C.c_int64; L.ora_self_kv_elements.argtypes = [vp]
        L.ora_get_self_kv.argtypes = [vp, fp, fp]
        L.ora_trace_name.restype = C.c_char_p; L.ora_trace_name.argtypes = [ci]
        L.ora_trace_shape.argtypes = [ci, C.POINTER(C.c_int64)]
        L.ora_trace_size.restype = C.c_int64; L.ora_trace_size.argtypes = [ci]
        L.ora_trace_data.argtypes = [ci, fp]
        L.ora_full.argtypes = [vp, fp, ci, ci, ci, C.c_char_p, ci, ci]
        L.ora_full_ex.argtypes = [vp, fp, ci, ci, ci, C.c_char_p, ci, ci, ci, ci]
        L.ora_full_ex2.argtypes = [vp, fp, ci, ci, ci, C.c_char_p, ci, ci, ci, ci]
        L.ora_full_token_t0.restype = C.c_int64; L.ora_full_token_t0.argtypes = [vp, ci, ci]
        L.ora_full_token_t1.restype = C.c_int64; L.ora_full_token_t1.argtypes = [vp, ci, ci]
        L.ora_full_token_vlen.restype = C.c_float; L.ora_full_token_vlen.argtypes = [vp, ci, ci]
        L.ora_full_n_segments.argtypes = [vp]
        L.ora_full_segment_t0.restype = C.c_int64; L.ora_full_segment_t0.argtypes = [vp, ci]
        L.ora_full_segment_t1.restype = C.c_int64; L.ora_full_segment_t1.argtypes = [vp, ci]
        L.ora_full_segment_text.restype = C.c_char_p; L.ora_full_segment_text.argtypes = [vp, ci]
        L.ora_full_n_tokens.argtypes = [vp, ci]
        L.ora_full_token_id.argtypes = [vp, ci, ci]
        L.ora_full_token_p.restype = C.c_float; L.ora_full_token_p.argtypes = [vp, ci, ci]
        L.ora_clear_prompt_past.argtypes = [vp]
        L.ora_bench_chunk.restype = C.c_double
        L.ora_bench_chunk.argtypes = [vp, fp, ci, ci, ip, ci, ci, ip, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class RefOracle:
    """One whisper_context of the reference CPU implementation."""

    def __init__(self, model_path: str, threads: int = 1, log_level: int = 1):
        self.L = lib()
        self.L.ora_set_log_level(log_level)
        self.ctx = self.L.ora_init(model_path.encode())
        if not self.ctx:
            raise RuntimeError("reference oracle failed to load " + model_path)
        self.threads = threads
        hp = np.zeros(11, np.int32)
        self.L.ora_hparams(self.ctx, _i(hp))
        (self.n_vocab, self.n_audio_ctx, self.n_audio_state, self.n_audio_head, self.n_audio_layer,
         self.n_text_ctx, self.n_text_state, self.n_text_head, self.n_text_layer, self.n_mels, self.f16) = [int(x) for x in hp]
        st = np.zeros(8, np.int32)
        self.L.ora_special_tokens(self.ctx, _i(st))
        self.special = dict(zip(["eot", "sot", "prev", "solm", "not", "beg", "translate", "transcribe"], [int(x) for x in st]))

    def close(self):
        if self.ctx:
            self.L.ora_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- stages -------------------------------------------------------------------------------
    def pcm_to_mel(self, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, np.float32)
        rc = self.L.ora_pcm_to_mel(self.ctx, _f(pcm), pcm.size, self.threads)
        assert rc == 0
        return self.get_mel()

    def get_mel(self) -> np.ndarray:
        n = self.L.ora_mel_len(self.ctx)
        out = np.empty((self.n_mels, n), np.float32)
        self.L.ora_get_mel(self.ctx, _f(out))
        return out

    def set_mel(self, mel: np.ndarray):
        mel = np.ascontiguousarray(mel, np.float32)
        assert self.L.ora_set_mel(self.ctx, _f(mel), mel.shape[1], mel.shape[0]) == 0

    def encode(self, offset: int = 0):
        assert self.L.ora_encode(self.ctx, offset, self.threads) == 0

    def decode(self, tokens, n_past: int):
        t = np.ascontiguousarray(tokens, np.int32)
        assert self.L.ora_decode(self.ctx, _i(t), t.size, n_past, self.threads) == 0
        n = self.L.ora_logits_size(self.ctx)
        logits = np.empty(n, np.float32)
        probs = np.empty(n, np.float32)
        self.L.ora_get_logits(self.ctx, _f(logits))
        self.L.ora_get_probs(self.ctx, _f(probs))
        return logits.reshape(t.size, self.n_vocab), probs.reshape(t.size, self.n_vocab)

    def tokenize(self, text: bytes, cap: int = 256):
        out = np.zeros(cap, np.int32)
        n = self.L.ora_tokenize(self.ctx, text, _i(out), cap)
        return None if n < 0 else out[:n].tolist()

    def sample(self, initial: bool = False, force_timestamp: bool = False):
        ids = np.zeros(2, np.int32)
        f3 = np.zeros(3, np.float32)
        if force_timestamp:
            self.L.ora_sample_timestamp(self.ctx, int(initial), _i(ids), _f(f3))
        else:
            self.L.ora_sample_best(self.ctx, _i(ids), _f(f3))
        return dict(id=int(ids[0]), tid=int(ids[1]), p=float(f3[0]), pt=float(f3[1]), ptsum=float(f3[2]))

    def sample_from_probs(self, probs: np.ndarray, force_timestamp: bool = False, is_initial: bool = False):
        """whisper_sample_best / whisper_sample_timestamp of the reference on a caller-supplied probability row"""
        p = np.ascontiguousarray(probs, np.float32)
        assert p.size == self.n_vocab
        ids = np.zeros(2, np.int32)
        f3 = np.zeros(3, np.float32)
        self.L.ora_sample_from_probs(self.ctx, _f(p), int(force_timestamp), int(is_initial), _i(ids), _f(f3))
        return dict(id=int(ids[0]), tid=int(ids[1]), p=float(f3[0]), pt=float(f3[1]), ptsum=float(f3[2]))

    def lang_auto_detect(self, offset_ms: int = 0):
        n = self.L.ora_lang_max_id() + 1
        pr = np.zeros(n, np.float32)
        lid = self.L.ora_lang_auto_detect(self.ctx, offset_ms, self.threads, _f(pr))
        return lid, pr

    def gap_log(self, on: bool | None = None):
        """on=True/False: start / stop recording the top-2 margin (logit units) of every whisper_decode call, also inside
        whisper_full; on=None: return what was recorded."""
        if on is not None:
            self.L.ora_gap_log_enable(self.ctx, int(on))
            return None
        n = self.L.ora_gap_log_count()
        out = np.zeros(n, np.float32)
        if n:
            self.L.ora_gap_log_get(_f(out))
        return out

    def cross_kv(self):
        n = self.L.ora_cross_kv_elements(self.ctx)
        k = np.empty(n, np.float32); v = np.empty(n, np.float32)
        self.L.ora_get_cross_kv(self.ctx, _f(k), _f(v))
        shp = (self.n_text_layer, self.n_audio_ctx, self.n_text_state)
        return k.reshape(shp), v.reshape(shp)

    def self_kv(self):
        n = self.L.ora_self_kv_elements(self.ctx)
        k = np.empty(n, np.float32); v = np.empty(n, np.float32)
        self.L.ora_get_self_kv(self.ctx, _f(k), _f(v))
        shp = (self.n_text_layer, self.n_text_ctx, self.n_text_state)
        return k.reshape(shp), v.reshape(shp)

    # -- tracing ------------------------------------------------------------------------------
    def trace(self, on: bool):
        self.L.ora_trace_enable(int(on))

    def trace_items(self):
        """dict name -> ndarray shaped (ne3, ne2, ne1, ne0) i.e. numpy order, last axis fastest.
        Repeated names get a #k suffix."""
        out = {}
        for i in range(self.L.ora_trace_count()):
            name = self.L.ora_trace_name(i).decode()
            ne = (C.c_int64 * 4)()
            self.L.ora_trace_shape(i, ne)
            n = self.L.ora_trace_size(i)
            a = np.empty(n, np.float32)
            self.L.ora_trace_data(i, _f(a))
            shape = tuple(int(x) for x in reversed(list(ne)))
            if int(np.prod(shape)) == n:
                a = a.reshape(shape)
            key, k = name, 1
            while key in out:
                key = "%s#%d" % (name, k); k += 1
            out[key] = a
        return out

    # -- whisper_full -------------------------------------------------------------------------
    def full(self, pcm, flags: int = 0, language: str = "en", max_tokens: int = 0, audio_ctx: int = 0):
        pcm = np.ascontiguousarray(pcm, np.float32)
        rc = self.L.ora_full(self.ctx, _f(pcm), pcm.size, self.threads, flags, language.encode(), max_tokens, audio_ctx)
        segs = []
        for i in range(self.L.ora_full_n_segments(self.ctx)):
            toks = [self.L.ora_full_token_id(self.ctx, i, j) for j in range(self.L.ora_full_n_tokens(self.ctx, i))]
            segs.append(dict(t0=self.L.ora_full_segment_t0(self.ctx, i), t1=self.L.ora_full_segment_t1(self.ctx, i),
                             text=self.L.ora_full_segment_text(self.ctx, i).decode(errors="replace"), tokens=toks))
        return rc, segs

    def clear_prompt_past(self):
        self.L.ora_clear_prompt_past(self.ctx)

    def bench_chunk(self, pcm, prompt, n_decode: int, threads: int | None = None):
        pcm = np.ascontiguousarray(pcm, np.float32)
        pr = np.ascontiguousarray(prompt, np.int32)
        toks = np.zeros(n_decode, np.int32)
        st = (C.c_double * 3)()
        s = self.L.ora_bench_chunk(self.ctx, _f(pcm), pcm.size, threads or self.threads, _i(pr), pr.size, n_decode, _i(toks), st)
        return float(s), toks, [float(x) for x in st]
