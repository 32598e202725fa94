#!/usr/bin/env python
"""GPU bring-up diagnostics: every stage runs in its own process under a timeout, so a trapping or hanging kernel in one
stage cannot take the others (or the GPU lease) with it.  Prints one line per check; used under gpurun:

    python tools/gpu_bringup.py            # all stages
    python tools/gpu_bringup.py gemm attn  # selected stages
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def err_stats(name, got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(got - ref)
    scale = np.maximum(1.0, np.abs(ref))
    print("  %-34s max_abs=%.3e  max_rel1=%.3e  rms=%.3e  ref_rms=%.3e  nan=%d" % (
        name, d.max(), (d / scale).max(), np.sqrt((d ** 2).mean()), np.sqrt((ref ** 2).mean()), int(np.isnan(got).sum())), flush=True)
    return d.max()


def stage_gemm():
    from whisper_b200 import capi
    rng = np.random.default_rng(0)
    for (M, N, K, bn) in [(128, 128, 64, 128), (128, 256, 128, 256), (300, 384, 200, 128), (1500, 1024, 1024, 256), (3000, 384, 384, 128), (777, 640, 72, 128)]:
        A = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
        B = (rng.standard_normal((N, K)) * 0.5).astype(np.float16)
        D, _ = capi.test_gemm(A, B, bn=bn)
        ref = A.astype(np.float32) @ B.astype(np.float32).T
        err_stats("gemm %dx%dx%d bn%d" % (M, N, K, bn), D, ref)


def stage_gemm_stress():
    """Repeat the large GEMM shapes and report WHERE any mismatch is (race hunting)."""
    from whisper_b200 import capi
    for (M, N, K, bn) in [(12000, 1280, 1280, 256), (12000, 1024, 1024, 256), (12000, 4096, 1024, 256), (12000, 1024, 4096, 256)]:
        rng = np.random.default_rng(M * 7 + N)
        A = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
        B = (rng.standard_normal((N, K)) * 0.5).astype(np.float16)
        ref = A.astype(np.float32) @ B.astype(np.float32).T
        bad = 0
        for it in range(40):
            D, _ = capi.test_gemm(A, B, bn=bn, iters=(3 if it % 2 else 0))
            err = np.abs(D - ref)
            if err.max() > 1e-3:
                bad += 1
                rows, cols = np.nonzero(err > 1e-3)
                print("  iter %d: %d bad elements, max err %.3e; rows %d..%d (tiles %s) cols %d..%d (tiles %s); sample D=%.4f ref=%.4f" % (
                    it, rows.size, err.max(), rows.min(), rows.max(), sorted(set((rows // 128).tolist()))[:8], cols.min(), cols.max(),
                    sorted(set((cols // bn).tolist()))[:8], D[rows[0], cols[0]], ref[rows[0], cols[0]]), flush=True)
        print("  gemm %dx%dx%d bn%d: %d / 40 runs with mismatches" % (M, N, K, bn, bad), flush=True)


def stage_gemm_perf():
    from whisper_b200 import capi
    rng = np.random.default_rng(1)
    for (M, N, K, bn) in [(12000, 1024, 1024, 256), (12000, 4096, 1024, 256), (12000, 1024, 4096, 256), (12000, 3072, 1024, 256), (12000, 1024, 1024, 128)]:
        A = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
        B = (rng.standard_normal((N, K)) * 0.5).astype(np.float16)
        D, ms = capi.test_gemm(A, B, bn=bn, iters=10)
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        print("  gemm %dx%dx%d bn%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, bn, ms, tf), flush=True)


def stage_ln():
    from whisper_b200 import capi
    rng = np.random.default_rng(2)
    for d in (128, 384, 1024, 1280):
        x = rng.standard_normal((37, d)).astype(np.float32) * 2 + 0.3
        g = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
        b = (0.1 * rng.standard_normal(d)).astype(np.float32)
        out = capi.test_layernorm(x, g, b)
        xd = x.astype(np.float64)
        ref = (xd - xd.mean(1, keepdims=True)) / np.sqrt(xd.var(1, keepdims=True) + 1e-5) * g + b
        err_stats("layernorm d=%d" % d, out, ref.astype(np.float16).astype(np.float32))


def stage_skinny():
    from whisper_b200 import capi
    rng = np.random.default_rng(3)
    for (nOut, K, cols) in [(64, 128, 1), (384, 384, 8), (1024, 1024, 8), (1024, 4096, 8), (51864, 384, 3), (3072, 1024, 24)]:
        W = (rng.standard_normal((nOut, K)) / np.sqrt(K)).astype(np.float16)
        X = rng.standard_normal((cols, K)).astype(np.float16)
        out, _ = capi.test_skinny(W, X)
        ref = X.astype(np.float32) @ W.astype(np.float32).T
        err_stats("skinny %dx%d cols=%d" % (nOut, K, cols), out, ref)
    W = (rng.standard_normal((51865, 1024)) / 32).astype(np.float16)
    X = rng.standard_normal((8, 1024)).astype(np.float16)
    _, ms = capi.test_skinny(W, X, iters=20)
    print("  skinny logits 51865x1024 cols=8: %.3f ms  %.0f GB/s" % (ms, W.nbytes / ms / 1e6), flush=True)
    W = (rng.standard_normal((4096, 1024)) / 32).astype(np.float16)
    _, ms = capi.test_skinny(W, X, iters=50)
    print("  skinny fc1 4096x1024 cols=8: %.4f ms  %.0f GB/s" % (ms, W.nbytes / ms / 1e6), flush=True)


def attn_ref(Q, K, V):
    Qf, Kf, Vf = Q.astype(np.float32), K.astype(np.float32), V.astype(np.float32)
    S = np.einsum("bqd,bkd->bqk", Qf, Kf) * 0.125
    S = S - S.max(-1, keepdims=True)
    P = np.exp(S)
    P = P / P.sum(-1, keepdims=True)
    P = P.astype(np.float16).astype(np.float32)
    return np.einsum("bqk,bkd->bqd", P, Vf)


def stage_attn():
    from whisper_b200 import capi
    rng = np.random.default_rng(4)
    for (BH, T) in [(1, 128), (2, 200), (3, 1500)]:
        Q = rng.standard_normal((BH, T, 64)).astype(np.float16)
        K = rng.standard_normal((BH, T, 64)).astype(np.float16)
        V = rng.standard_normal((BH, T, 64)).astype(np.float16)
        out, _ = capi.test_attention(Q, K, V)
        err_stats("attention BH=%d T=%d" % (BH, T), out, attn_ref(Q, K, V).astype(np.float16).astype(np.float32))
    BH, T = 128, 1500
    Q = rng.standard_normal((BH, T, 64)).astype(np.float16)
    out, ms = capi.test_attention(Q, Q, Q, iters=10)
    print("  attention BH=128 T=1500: %.3f ms  %.1f TFLOP/s" % (ms, 4.0 * BH * T * T * 64 / (ms * 1e-3) / 1e12), flush=True)


def _open(model_name, batch=1):
    from whisper_b200 import capi, synth
    path = synth.model_path(model_name)
    m = capi.Model(path)
    e = capi.Engine(m, 0)
    c = capi.Context(e, batch)
    return path, m, e, c


def stage_mel():
    from whisper_b200 import synth
    from oracle.ref import RefOracle
    path, m, e, c = _open("micro.en")
    o = RefOracle(path, threads=4)
    for cid, n in [(0, 480000), (1, 176000), (2, 16000 * 7 + 123)]:
        pcm = synth.synth_pcm(cid, n)
        ref = o.pcm_to_mel(pcm)
        c.pcm_to_mel(0, pcm)
        got = c.get_mel(0)
        print("  mel shapes", got.shape, ref.shape, flush=True)
        err_stats("mel chunk %d n=%d" % (cid, n), got, ref)


def stage_encoder(model_name="micro.en"):
    from whisper_b200 import synth
    from oracle.ref import RefOracle
    path = synth.model_path(model_name)
    o = RefOracle(path, threads=int(os.environ.get("ORA_THREADS", "4")))
    pcm = synth.synth_pcm(0)
    mel = o.pcm_to_mel(pcm)
    print("  oracle mel done", flush=True)
    o.trace(True)
    o.encode(0)
    print("  oracle encode done", flush=True)
    tr = o.trace_items()
    o.trace(False)
    print("  oracle trace keys:", list(tr.keys())[:12], flush=True)
    path, m, e, c = _open(model_name)
    c.set_mel(0, mel)
    d, T = m.n_audio_state, m.n_audio_ctx
    L = m.n_audio_layer
    # conv front end
    c.set_encoder_layers(0)
    c.encode(1)
    conv1 = c.get_tensor("enc.conv1").reshape(3000, d)
    err_stats("enc.temp1 (conv1+gelu)", conv1, tr["enc.temp1"].reshape(d, 3000).T)
    x0 = c.get_tensor("enc.x").reshape(T, d)
    err_stats("enc.layer[0].in", x0, tr["enc.layer[ 0 ].in"].reshape(T, d))
    for nl in range(1, L + 1):
        c.set_encoder_layers(nl)
        c.encode(1)
        x = c.get_tensor("enc.x").reshape(T, d)
        key = "enc.layer[ %d ].in" % nl if nl < L else "enc.layers"
        err_stats("after %d layers (%s)" % (nl, key), x, tr[key].reshape(T, d))
        if nl == 1:
            q = c.get_tensor("enc.q").reshape(m.n_audio_head, T, 64)
            qref = tr["enc-Qcur"].reshape(T, d) + 0  # before bias
            attn = c.get_tensor("enc.attn").reshape(T, m.n_audio_head, 64)
            kqv = tr["enc-KQV"].reshape(m.n_audio_head, T, 64).transpose(1, 0, 2)
            err_stats("layer0 attention out (enc-KQV)", attn, kqv)
    c.set_encoder_layers(-1)
    c.encode(1)
    out = c.get_tensor("encode-out").reshape(T, d)
    err_stats("encode-out", out, tr["encode-out"].reshape(T, d))
    ck, cv = o.cross_kv()
    gk = c.get_tensor("cross_k").reshape(m.n_text_layer, m.n_audio_head, T, 64).transpose(0, 2, 1, 3).reshape(m.n_text_layer, T, d)
    gv = c.get_tensor("cross_v").reshape(m.n_text_layer, m.n_audio_head, T, 64).transpose(0, 2, 1, 3).reshape(m.n_text_layer, T, d)
    err_stats("cross K", gk, ck)
    err_stats("cross V", gv, cv)
    ms, calls = c.timings()
    print("  timings ms", ms, calls, flush=True)


def stage_decoder(model_name="micro.en"):
    from whisper_b200 import capi, synth
    from oracle.ref import RefOracle
    path, m, e, c = _open(model_name)
    th = int(os.environ.get("ORA_THREADS", "4"))
    o = RefOracle(path, threads=th)
    c.set_reference_threads(th)
    print("  reference threads = %d" % th, flush=True)
    pcm = synth.synth_pcm(0)
    mel = o.pcm_to_mel(pcm)
    o.encode(0)
    c.set_mel(0, mel)
    c.encode(1)
    prompt = m.prompt_init()
    # teacher-forced: feed the oracle's tokens to both
    lg_ref, pr_ref = o.decode(prompt, 0)
    c.decode([prompt], 0, 1, capi.DECODE_ALL_LOGITS)
    lg = c.logits(len(prompt))
    err_stats("prompt logits (all rows)", lg, lg_ref)
    err_stats("prompt probs", c.probs(len(prompt)), pr_ref)
    tok_ref = o.sample(initial=True, force_timestamp=True)
    s = c.decode([prompt], 0, 1, capi.DECODE_FORCE_TIMESTAMP | capi.DECODE_INITIAL)[0]
    print("  sample initial: ref", tok_ref, "got", s, flush=True)
    n_past = len(prompt)
    cur = tok_ref["id"]
    nmatch = 0
    worst = 0.0
    for i in range(24):
        lg_ref, pr_ref = o.decode([cur], n_past)
        t_ref = o.sample()
        s = c.decode([[cur]], n_past, 1, 0)[0]
        lg = c.logits(1)
        dmax = np.abs(lg - lg_ref).max()
        worst = max(worst, dmax)
        srt = np.sort(lg_ref[0])
        nmatch += int(s["id"] == t_ref["id"])
        if i < 6 or s["id"] != t_ref["id"]:
            print("   step %d: ref id %d p %.4f | got id %d p %.4f | max|dlogit| %.3e top2gap %.3e" % (
                i, t_ref["id"], t_ref["p"], s["id"], s["p"], dmax, srt[-1] - srt[-2]), flush=True)
        n_past += 1
        cur = t_ref["id"]
    print("  teacher-forced 24 steps: token match %d/24, worst |dlogit| %.3e" % (nmatch, worst), flush=True)
    # graph on/off consistency + device-token feedback
    toks, st = c.run_chunks([pcm], prompt, 16)
    print("  run_chunks tokens:", toks[0].tolist(), "stage_ms", st.tolist(), flush=True)
    c.set_graph(False)
    toks2, st2 = c.run_chunks([pcm], prompt, 16)
    print("  run_chunks (no graph):", "same" if (toks2 == toks).all() else toks2[0].tolist(), st2.tolist(), flush=True)
    c.set_step_mode(0)
    toks3, st3 = c.run_chunks([pcm], prompt, 16)
    print("  run_chunks (no mega, no graph):", "same" if (toks3 == toks).all() else toks3[0].tolist(), st3.tolist(), flush=True)
    c.set_step_mode(2); c.set_graph(True)
    ref_s, ref_toks, ref_st = o.bench_chunk(pcm, prompt, 16, threads=th)
    print("  oracle tokens    :", ref_toks.tolist(), flush=True)


def stage_batch(model_name="micro.en"):
    from whisper_b200 import synth
    path, m, e, c = _open(model_name, batch=4)
    pcms = [synth.synth_pcm(i) for i in range(4)]
    prompt = m.prompt_init()
    toks, st = c.run_chunks(pcms, prompt, 12)
    print("  batch4 tokens:", toks.tolist(), st.tolist(), flush=True)
    c1 = __import__("whisper_b200.capi", fromlist=["Context"]).Context(e, 1)
    for i in range(4):
        t1, _ = c1.run_chunks([pcms[i]], prompt, 12)
        print("  single[%d]   :" % i, t1[0].tolist(), "match" if (t1[0] == toks[i]).all() else "MISMATCH", flush=True)


def stage_encoder_tiny():
    stage_encoder("tiny.en")


def stage_decoder_tiny():
    stage_decoder("tiny.en")


def stage_perf(model_name="medium", batch=8, n_decode=100):
    from whisper_b200 import synth
    t0 = time.time()
    path, m, e, c = _open(model_name, batch=batch)
    print("  load %s: %.1f s, weights %.1f MB" % (model_name, time.time() - t0, e.weight_bytes() / 1e6), flush=True)
    pcms = [synth.synth_pcm(i) for i in range(batch)]
    prompt = m.prompt_init()
    for it in range(3):
        t0 = time.time()
        toks, st = c.run_chunks(pcms, prompt, n_decode)
        dt = time.time() - t0
        print("  %s B=%d n_decode=%d: wall %.1f ms  stages(mel,enc,dec)=%s  audio-s/s=%.0f" % (
            model_name, batch, n_decode, dt * 1e3, np.round(st, 2).tolist(), 30.0 * batch / dt), flush=True)
    print("  tokens[0][:16]", toks[0][:16].tolist(), flush=True)


def stage_profile(model_name="medium", batch=8, n_decode=3):
    """short run for ncu: one mel + encode + n_decode steps, no graph"""
    from whisper_b200 import synth
    path, m, e, c = _open(model_name, batch=batch)
    c.set_graph(os.environ.get("WSP_GRAPH", "0") == "1")
    c.set_reference_threads(int(os.environ.get("FLOW_REFTHREADS", "4")))
    n_decode = int(os.environ.get("PROFILE_NDECODE", n_decode))
    pcms = [synth.synth_pcm(i) for i in range(batch)]
    toks, st = c.run_chunks(pcms, m.prompt_init(), n_decode)
    print("  profile run stages", st.tolist(), flush=True)


def stage_flow():
    """The dataflow decoder-step kernel against round 1's barrier kernel and the kernel-per-op path: same tokens, same logits (the GEMV
    and cross-attention arithmetic is bit-identical; self-attention scores are summed in a different order), and the decode time of each."""
    from whisper_b200 import capi, synth
    cases = [("micro.en", 1, 20), ("micro.en", 4, 20), ("tiny.en", 8, 40), ("base.en", 3, 40), ("medium", 8, 100), ("large", 2, 20), ("medium", 12, 24), ("medium-sc", 8, 60)]
    only = os.environ.get("FLOW_CASES")
    if only:
        cases = [cs for cs in cases if cs[0] in only.split(",")]
    ref_threads = int(os.environ.get("FLOW_REFTHREADS", "4"))
    for (name, batch, n) in cases:
        path, m, e, c = _open(name, batch=batch)
        c.set_reference_threads(ref_threads)
        pcms = [synth.synth_pcm(i) for i in range(batch)]
        prompt = m.prompt_init()
        res = {}
        for mode in [int(x) for x in os.environ.get("FLOW_MODES", "2,1,0").split(",")]:
            c.set_step_mode(mode)
            try:
                toks, st = c.run_chunks(pcms, prompt, n)
                toks, st = c.run_chunks(pcms, prompt, n)
                res[mode] = (toks.copy(), c.logits(batch).copy(), st[2])
            except Exception as ex:  # noqa: BLE001
                print("  %s B=%d mode %d FAILED: %s" % (name, batch, mode, ex), flush=True)
                break
        if 2 in res and 0 in res:
            for mode in (1, 0):
                if mode not in res:
                    continue
                same = (res[2][0] == res[mode][0]).all()
                dl = np.abs(res[2][1] - res[mode][1]).max()
                first = ""
                if not same:
                    d = np.argwhere(res[2][0] != res[mode][0])
                    first = " (first difference: chunk %d step %d; %d of %d chunks differ)" % (d[0][0], d[:, 1].min(), len(set(d[:, 0].tolist())), batch)
                print("  %-9s B=%-2d n=%-3d flow vs mode %d: tokens %s%s, max|dlogit| %.3e | decode ms flow %.2f  mode%d %.2f" % (
                    name, batch, n, mode, "same" if same else "DIFFER", first, dl, res[2][2], mode, res[mode][2]), flush=True)
            print("     tokens[0][:12]", res[2][0][0][:12].tolist(), flush=True)
        c.set_step_mode(2)
        c.close(); e.close(); m.close()


def stage_steptiming(model_name="medium", batch=8):
    """Where one decoder step goes, as CTA `WSP_TIMING_CTA` (default 0) of the dataflow kernel sees it: (id, %globaltimer) marks at every
    phase boundary and at the sub-steps of a phase; printed as the mean time from the PREVIOUS mark, averaged over the layers."""
    from whisper_b200 import synth
    model_name = os.environ.get("STEP_MODEL", model_name)
    batch = int(os.environ.get("STEP_BATCH", batch))
    path, m, e, c = _open(model_name, batch=batch)
    pcms = [synth.synth_pcm(i) for i in range(batch)]
    c.set_reference_threads(int(os.environ.get("FLOW_REFTHREADS", "4")))
    c.step_timing(True)
    c.run_chunks(pcms, m.prompt_init(), 40)
    rawall = c.step_timing().astype(np.int64).reshape(-1, 2)
    prod = rawall[2000:]
    prod = prod[prod[:, 1] != 0]
    raw = rawall[:2000]
    raw = raw[: int(np.nonzero(raw[:, 1])[0].max()) + 1] if raw[:, 1].any() else raw[:0]
    ids, t = raw[:, 0], raw[:, 1]
    L = m.n_text_layer
    names = ["LN1+QKV", "self-attn", "O+res", "LN+CQ", "cross-attn", "CO+res", "LN+FC1", "FC2+res"]
    subn = {0: "phase done", 1: "inputs staged", 2: "1st weights landed" , 3: "MMAs done", 4: "stored"}
    suba = {1: "q arrived", 2: "scores done", 3: "softmax done", 4: "chains done", 0: "phase done"}
    suba.update({10 + k: "K%d" % k for k in range(24)})
    suba.update({40 + k: "V%d" % k for k in range(24)})
    acc = {}
    for k in range(1, len(ids)):
        acc.setdefault(int(ids[k]), []).append((t[k] - t[k - 1]) / 1e3)
    print("  %s B=%d CTA %s: step %.1f us, %d marks" % (model_name, batch, os.environ.get("WSP_TIMING_CTA", "0"), (t[-1] - t[0]) / 1e3, len(ids)), flush=True)
    phase_tot = {}
    for key in sorted(acc):
        ph, sb = key // 100 - 1, key % 100
        lay, p8 = divmod(ph, 8)
        phase_tot.setdefault(p8 if ph < 8 * L else 8, []).append(acc[key])
    per_phase = {}
    for key in sorted(acc):
        ph, sb = key // 100 - 1, key % 100
        p8 = ph % 8 if ph < 8 * L else 8
        per_phase.setdefault((p8, sb), []).extend(acc[key])
    for p8 in range(9):
        items = sorted((sb if sb else 99, sb) for (pp, sb) in per_phase if pp == p8)
        if not items:
            continue
        nm = names[p8] if p8 < 8 else "logits"
        tot = sum(np.mean(per_phase[(p8, sb)]) for _, sb in items)
        lab = suba if p8 in (1, 4) else subn
        items = sorted(items, key=lambda it: (it[1] == 0, it[1]))
        print("    %-10s %6.2f us | " % (nm, tot) + "  ".join("%s %.2f" % (lab.get(sb, "#%d" % sb), np.mean(per_phase[(p8, sb)])) for _, sb in items), flush=True)
    _producer_report(prod, ids, t, L)


def _producer_report(prod, ids, t, L):
    """When the producer warp ISSUED selected loads of a layer, relative to the consumers' phase boundaries of the same layer."""
    if not len(prod):
        return
    cons = {int(i): int(tt) for i, tt in zip(ids, t)}
    names = {1: "Wo", 2: "Pc", 3: "crossK0", 4: "crossK4", 5: "crossK8", 6: "crossV0", 7: "after crossV", 8: "after Wco"}
    # consumer marks: phase p of layer l ends at id 100 * (8 l + p + 1); cross-attention starts when CQ (p = 3) ends
    rel = {}
    for pid, pt in prod:
        il, code = divmod(int(pid) - 100000, 100)
        if il < 1 or il >= L - 1:
            continue
        for nm, ph in (("self-attn end", 1), ("cross-attn start", 3), ("cross-attn end", 4)):
            key = 100 * (8 * il + ph + 1)
            if key in cons:
                rel.setdefault((code, nm), []).append((pt - cons[key]) / 1e3)
    for code in sorted(names):
        parts = ["%s %+.2f" % (nm, np.mean(rel[(code, nm)])) for nm in ("self-attn end", "cross-attn start", "cross-attn end") if (code, nm) in rel]
        if parts:
            print("    producer issues %-13s at (us relative to the consumers'): %s" % (names[code], "  ".join(parts)), flush=True)


STAGES = {
    "gemm": stage_gemm, "ln": stage_ln, "skinny": stage_skinny, "attn": stage_attn, "mel": stage_mel,
    "encoder": stage_encoder, "decoder": stage_decoder, "batch": stage_batch,
    "encoder_tiny": stage_encoder_tiny, "decoder_tiny": stage_decoder_tiny,
    "gemm_perf": stage_gemm_perf, "gemm_stress": stage_gemm_stress, "perf": stage_perf, "profile": stage_profile, "flow": stage_flow, "steptiming": stage_steptiming,
}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--stage":
        import faulthandler
        faulthandler.enable()
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES.keys())
    for n in names:
        print("=== stage %s ===" % n, flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", n], timeout=420)
            print("=== stage %s exit %d (%.1f s) ===" % (n, r.returncode, time.time() - t0), flush=True)
        except subprocess.TimeoutExpired:
            print("=== stage %s TIMEOUT ===" % n, flush=True)
