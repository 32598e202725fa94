#!/usr/bin/env python
"""How far do two legitimate builds of the REFERENCE differ from each other?  Runs the bench loop (timestamp-first sampling, then
whisper_sample_best) of oracle/_ref (AVX2+FMA+F16C, the build the fixtures come from) and of the scalar build of the same
sources (`make -C oracle scalar`) on the same model / chunk and compares the last step's logits.  Test infrastructure only.

usage: python tools/reference_spread.py MODEL [CHUNK] [STEPS]      (one process per build: the library path is fixed at import)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(model, chunk, steps, out):
    import ctypes as C
    import numpy as np
    from oracle.ref import RefOracle
    from whisper_b200 import synth
    o = RefOracle(synth.model_path(model), threads=4)
    prompt = np.array([o.special["sot"]] + ([o.special["sot"] + 1, o.special["transcribe"]] if o.n_vocab == 51865 else []), np.int32)
    secs, toks, st = o.bench_chunk(synth.synth_pcm(chunk), prompt, steps, threads=4)
    logits = np.empty(o.L.ora_logits_size(o.ctx), np.float32)
    o.L.ora_get_logits(o.ctx, logits.ctypes.data_as(C.POINTER(C.c_float)))
    np.savez(out, tokens=toks, logits=logits, secs=secs)


def main():
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
        return
    import numpy as np
    model = sys.argv[1]
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    res = {}
    for name, lib in (("avx2", "liboracle_ref.so"), ("scalar", "liboracle_ref_scalar.so")):
        out = "/tmp/spread_%s_%s_%d.npz" % (name, model, chunk)
        env = dict(os.environ, WSP_ORACLE_LIB=os.path.join(ROOT, "oracle", "_ref", lib))
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", model, str(chunk), str(steps), out], check=True, env=env)
        res[name] = np.load(out)
    a, b = res["avx2"], res["scalar"]
    d = np.abs(a["logits"] - b["logits"])
    print(json.dumps({"model": model, "chunk": chunk, "steps": steps, "tokens_equal": bool((a["tokens"] == b["tokens"]).all()),
                      "logit_max_abs_diff": float(d.max()), "logit_rms_diff": float(np.sqrt((d * d).mean())),
                      "logit_rms": float(np.sqrt((a["logits"] ** 2).mean())), "seconds": [float(a["secs"]), float(b["secs"])]}))


if __name__ == "__main__":
    main()
