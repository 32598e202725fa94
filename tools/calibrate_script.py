#!/usr/bin/env python
"""TEST INFRASTRUCTURE — calibration of the scripted synthetic models (whisper_b200/synth.py, names ending in "-sc").

A random encoder / decoder emits ~95 % one constant vector.  The scripted models subtract that constant (encoder: from the ln_post
output, decoder: from the final LayerNorm output) and amplify the remainder, so that greedy decoding really depends on the audio
and on the history.  The constants are measured here with the reference's own CPU code (oracle/_ref) and stored in
whisper_b200/script_calib.npz; synth.py then writes the final model files deterministically, with or without the oracle present.

    python tools/calibrate_script.py micro.en micro tiny.en base.en medium large
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import whisper_np as wn  # noqa: E402
from oracle.ref import RefOracle  # noqa: E402
from whisper_b200 import synth  # noqa: E402

CHUNKS = (0, 5)
N_STEPS = 60
SEED = 1234


def prompt_of(o):
    return [o.special["sot"]] + ([o.special["sot"] + 1, o.special["transcribe"]] if o.n_vocab == 51865 else [])


def calibrate(base: str, tmp_dir: str = "/tmp/wsp_models"):
    os.makedirs(tmp_dir, exist_ok=True)
    p = os.path.join(tmp_dir, "calib-%s.bin" % base)
    # pass 1: encoder output without centring -> its mean over time and chunks, at unit gain
    synth.write_script_model(p, base, SEED, stage=0)
    o = RefOracle(p, threads=8)
    m = wn.NpModel(p)
    g, b = m.t["encoder.ln_post.weight"].astype(np.float64), m.t["encoder.ln_post.bias"].astype(np.float64)
    outs = []
    for ch in CHUNKS:
        o.pcm_to_mel(synth.synth_pcm(ch))
        o.trace(True); o.encode(0); tr = o.trace_items(); o.trace(False)
        outs.append((tr["encode-out"].reshape(o.n_audio_ctx, o.n_audio_state).astype(np.float64) - b) / g)
    enc = np.concatenate(outs).mean(0)
    dev = (np.concatenate(outs) - enc).std()
    o.close()
    # pass 2: decoder's final LayerNorm output (recovered from the logits through the embedding matrix), encoder centred
    synth.write_script_model(p, base, SEED, stage=1, calib={"enc": enc, "dec": None})
    o = RefOracle(p, threads=8)
    m = wn.NpModel(p)
    E = m.t["decoder.token_embedding.weight"].astype(np.float64)
    rows = []
    for ch in CHUNKS:
        o.pcm_to_mel(synth.synth_pcm(ch)); o.encode(0)
        pr = prompt_of(o)
        lg, _ = o.decode(pr, 0)
        rows.append(lg[-1].astype(np.float64))
        tok = o.sample(initial=True, force_timestamp=True)["id"]
        n_past = len(pr)
        for _ in range(N_STEPS - 1):
            lg, _ = o.decode([tok], n_past); n_past += 1
            rows.append(lg[0].astype(np.float64))
            tok = o.sample()["id"]
    o.close()
    Z = np.linalg.lstsq(E, np.array(rows).T, rcond=None)[0].T / m.t["decoder.ln.weight"].astype(np.float64)
    dec = Z.mean(0)
    var = float(np.sqrt((((Z - dec)[:, synth.SC_NC:]) ** 2).sum(1).mean()))
    print("%-8s encoder |c| %.2f deviation rms %.3f | decoder noise |c| %.2f variable part %.2f" % (
        base, np.linalg.norm(enc), dev, np.linalg.norm(dec[synth.SC_NC:]), var), flush=True)
    os.remove(p)
    return enc.astype(np.float32), dec.astype(np.float32), np.float32(var)


if __name__ == "__main__":
    names = sys.argv[1:] or ["micro.en", "micro", "tiny.en", "base.en", "medium", "large"]
    data = dict(np.load(synth.CALIB_PATH)) if os.path.exists(synth.CALIB_PATH) else {}
    for nm in names:
        enc, dec, var = calibrate(nm)
        data["%s_%d_enc" % (nm, SEED)] = enc
        data["%s_%d_dec" % (nm, SEED)] = dec
        data["%s_%d_var" % (nm, SEED)] = var
        np.savez_compressed(synth.CALIB_PATH, **data)
    print("wrote", synth.CALIB_PATH, "%.0f KB" % (os.path.getsize(synth.CALIB_PATH) / 1024))
