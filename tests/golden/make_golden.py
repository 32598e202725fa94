#!/usr/bin/env python
"""Generate tests/golden/*.npz from oracle/_ref — the reference's own CPU implementation (Whisper/source/{ggml.c,whisper.cpp},
compiled unmodified by oracle/Makefile).  The reference ships no golden vectors for this path, so these
fixtures ARE the pin: they are what the numpy restatement (oracle/whisper_np.py) and the CUDA engine are compared with on
machines where /root/reference (and hence a fresh oracle/_ref build) is not available.

Inputs are fully determined by whisper_b200/synth.py (model seed 1234, PCM chunk ids), so only outputs are stored, and large
tensors are stored as deterministic sub-samples to keep the fixtures small.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref import RefOracle  # noqa: E402
from whisper_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# deterministic sub-sampling shared with the tests
MEL_STEP = 10        # every 10th frame
ROW_STEP = 25        # every 25th time step of [T][d] tensors
LOGIT_STEP = 61      # every 61st vocabulary entry
N_STEPS = 16         # greedy steps stored


def logits_sample(lg):
    idx = np.arange(0, lg.shape[-1], LOGIT_STEP)
    return lg[..., idx].astype(np.float32)


def make(model_name: str, chunk: int, n_samples: int, offset: int, threads_list=(1, 4)):
    path = synth.model_path(model_name)
    pcm = synth.synth_pcm(chunk, n_samples)
    out = {}
    o = RefOracle(path, threads=1)
    mel = o.pcm_to_mel(pcm)
    out["mel_shape"] = np.array(mel.shape, np.int32)
    out["mel"] = mel[:, ::MEL_STEP].astype(np.float32)
    o.trace(True)
    o.encode(offset)
    tr = o.trace_items()
    o.trace(False)
    d, T, L = o.n_audio_state, o.n_audio_ctx, o.n_audio_layer
    out["enc_temp1"] = tr["enc.temp1"].reshape(d, 3000).T[::ROW_STEP * 2].astype(np.float32)
    for il in (0, 1):
        out["enc_layer%d_in" % il] = tr["enc.layer[ %d ].in" % il].reshape(T, d)[::ROW_STEP].astype(np.float32)
    out["enc_layers"] = tr["enc.layers"].reshape(T, d)[::ROW_STEP].astype(np.float32)
    out["encode_out"] = tr["encode-out"].reshape(T, d)[::ROW_STEP].astype(np.float32)
    ck, cv = o.cross_kv()
    out["cross_k"] = ck[:, ::ROW_STEP].astype(np.float16)
    out["cross_v"] = cv[:, ::ROW_STEP].astype(np.float16)
    prompt = [o.special["sot"]]
    if o.n_vocab == 51865:
        prompt += [o.special["sot"] + 1, o.special["transcribe"]]
    out["prompt"] = np.array(prompt, np.int32)
    for th in threads_list:
        oo = RefOracle(path, threads=th)
        oo.set_mel(mel)
        oo.encode(offset)
        lg, pr = oo.decode(prompt, 0)
        out["t%d_prompt_logits" % th] = logits_sample(lg)
        out["t%d_prompt_logits_max" % th] = lg.max(-1).astype(np.float32)
        out["t%d_prompt_argmax" % th] = lg.argmax(-1).astype(np.int32)
        tok = oo.sample(initial=True, force_timestamp=True)
        toks, ps, step_logits, tids = [tok["id"]], [tok["p"]], [], [tok["tid"]]
        n_past = len(prompt)
        for _ in range(N_STEPS - 1):
            lg, pr = oo.decode([toks[-1]], n_past)
            n_past += 1
            step_logits.append(logits_sample(lg[0]))
            tok = oo.sample()
            toks.append(tok["id"]); ps.append(tok["p"]); tids.append(tok["tid"])
        out["t%d_tokens" % th] = np.array(toks, np.int32)
        out["t%d_tids" % th] = np.array(tids, np.int32)
        out["t%d_token_p" % th] = np.array(ps, np.float32)
        out["t%d_step_logits" % th] = np.stack(step_logits)
    return out


CASES = {
    # name: (model, pcm chunk id, n_samples, mel offset)
    "micro_en_30s": ("micro.en", 0, 480000, 0),
    "micro_ml_11s": ("micro", 1, 176000, 0),          # multilingual specials, short clip (zero-padded window)
    "micro_en_offset": ("micro.en", 2, 640000, 1000),  # 40 s clip, window starting at frame 1000 (ragged tail)
}

# whisper_full() runs (the transcription driver, whisper.cpp:2765-3125) on a timestamp-happy synthetic model
FULL_MODEL = "micro.en-ts"
FULL_SECONDS = 75
FULL_RUNS = {
    # name: (eFullParamsFlags, max_tokens, offset_ms, duration_ms)      flags: 2 NoContext, 4 SingleSegment, 8 PrintSpecial
    "plain": (0, 0, 0, 0),
    "special": (8, 0, 0, 0),
    "special_nocontext_max40": (8 | 2, 40, 0, 0),
    "single_segment_max16": (4, 16, 0, 0),
    "special_offset_duration": (8 | 2, 30, 10000, 40000),
}


def full_pcm():
    n = 16000 * FULL_SECONDS
    return np.concatenate([synth.synth_pcm(10 + i) for i in range((n + 479999) // 480000)])[:n]


def make_full():
    path = synth.model_path(FULL_MODEL)
    pcm = full_pcm()
    out = {}
    for name, (flags, max_tokens, off, dur) in FULL_RUNS.items():
        o = RefOracle(path, threads=4, log_level=0)
        L = o.L
        import ctypes as C
        # ora_full has no offset/duration arguments: emulate them by slicing is NOT equivalent, so extend through the params struct
        rc = L.ora_full_ex(o.ctx, pcm.ctypes.data_as(C.POINTER(C.c_float)), pcm.size, 4, flags, b"en", max_tokens, 0, off, dur)
        assert rc == 0
        segs = []
        for i in range(L.ora_full_n_segments(o.ctx)):
            toks = [L.ora_full_token_id(o.ctx, i, j) for j in range(L.ora_full_n_tokens(o.ctx, i))]
            segs.append((L.ora_full_segment_t0(o.ctx, i), L.ora_full_segment_t1(o.ctx, i), toks, L.ora_full_segment_text(o.ctx, i)))
        out[name + "_t"] = np.array([[s[0], s[1]] for s in segs], np.int64).reshape(-1, 2)
        out[name + "_ntok"] = np.array([len(s[2]) for s in segs], np.int32)
        out[name + "_tokens"] = np.array([t for s in segs for t in s[2]], np.int32)
        out[name + "_text"] = np.array([s[3].decode(errors="replace") for s in segs])
        print("  full/%s: %d segments, %d tokens" % (name, len(segs), out[name + "_tokens"].size))
    return out


# Real model shapes (BASELINE.json configs): greedy tokens of the bench loop (timestamp-first sampling, then whisper_sample_best) at
# 4 reference threads, plus a subsample of the last step's logits.  Small files; the reference needs seconds (tiny) to a minute (medium).
REAL_SHAPES = {"tiny.en": [0, 5], "base.en": [0, 5], "medium": [0, 5]}
REAL_STEPS = 24


def make_real_shapes():
    out = {}
    for model, chunks in REAL_SHAPES.items():
        o = RefOracle(synth.model_path(model), threads=4)
        prompt = np.array([o.special["sot"]] + ([o.special["sot"] + 1, o.special["transcribe"]] if o.n_vocab == 51865 else []), np.int32)
        for ch in chunks:
            pcm = synth.synth_pcm(ch)
            secs, toks, st = o.bench_chunk(pcm, prompt, REAL_STEPS, threads=4)   # = the arithmetic the fixtures are compared under
            logits = np.empty(o.L.ora_logits_size(o.ctx), np.float32)
            o.L.ora_get_logits(o.ctx, logits.ctypes.data_as(C.POINTER(C.c_float)))
            key = "%s_c%d" % (model.replace(".", "_"), ch)
            out[key + "_tokens"] = toks.astype(np.int32)
            out[key + "_prompt"] = prompt
            out[key + "_last_logits_sub"] = logits[::LOGIT_STEP].astype(np.float32)
            print("  real/%s chunk %d: %.1f s, tokens %s..." % (model, ch, secs, toks[:6].tolist()), flush=True)
        del o
    return out


TOKENIZER_TEXTS = [b" hello hellox 12345!", b" hello world's , worlds 345 12", b" t12 t345", b""]


def make_tokenizer():
    """Prints the TOKENIZER_GOLDEN table of tests/test_gpu_com.py: whisper_tokenize of the reference on the "-words" vocabulary."""
    o = RefOracle(synth.model_path("micro.en-words"))
    for t in TOKENIZER_TEXTS:
        print("    %r: %r," % (t, o.tokenize(t)))


if __name__ == "__main__":
    if "--tokenizer" in sys.argv:
        make_tokenizer()
        sys.exit(0)
    if "--real-shapes" in sys.argv:
        data = make_real_shapes()
        p = os.path.join(HERE, "real_shapes.npz")
        np.savez_compressed(p, **data)
        print("real_shapes", "%.0f KB" % (os.path.getsize(p) / 1024))
        sys.exit(0)
    data = make_full()
    p = os.path.join(HERE, "full_micro_en_ts.npz")
    np.savez_compressed(p, **data)
    print("full_micro_en_ts", "%.0f KB" % (os.path.getsize(p) / 1024))
    for name, (model, chunk, n, off) in CASES.items():
        data = make(model, chunk, n, off)
        p = os.path.join(HERE, name + ".npz")
        np.savez_compressed(p, **data)
        print(name, "%.0f KB" % (os.path.getsize(p) / 1024))
