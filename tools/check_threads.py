#!/usr/bin/env python
"""Which side is right when the dataflow step kernel and the per-op kernels disagree on a fragile (unscripted) model: run the reference
itself (oracle/_ref) at the same reference thread count on the same chunk and print where each GPU path leaves its greedy sequence, with
the reference's top-2 logit margin at that step.  usage (GPU box): python tools/check_threads.py [model] [chunk] [threads] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.ref import RefOracle                      # noqa: E402
from tests.golden.make_golden import greedy, prompt_of  # noqa: E402
from whisper_b200 import capi, synth                  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "medium"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 6
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 16
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100

path = synth.model_path(model)
pcm = synth.synth_pcm(chunk)
o = RefOracle(path, threads=threads)
o.pcm_to_mel(pcm)
o.encode(0)
prompt = prompt_of(o)
r = greedy(o, prompt, steps)
ref = r["tokens"]
print("reference (%d threads) tokens[:12] %s  min gap %.4f" % (threads, ref[:12].tolist(), r["gap"].min()), flush=True)

m = capi.Model(path)
e = capi.Engine(m, 0)
c = capi.Context(e, 1)
c.set_reference_threads(threads)
for mode in (2, 1, 0):
    c.set_step_mode(mode)
    toks, _ = c.run_chunks([pcm], prompt, steps)
    t = toks[0][:steps]
    d = np.nonzero(t != ref)[0]
    if len(d):
        k = int(d[0])
        print("mode %d: leaves the reference at step %d (reference margin there %.4f; reference %d, GPU %d)" % (
            mode, k, r["gap"][k - 1] if k > 0 else float("nan"), ref[k], t[k]), flush=True)
    else:
        print("mode %d: identical to the reference for %d steps" % (mode, steps), flush=True)
