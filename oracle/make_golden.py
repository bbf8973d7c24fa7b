"""Generate tests/golden/* from the UNMODIFIED reference (TEST INFRASTRUCTURE — see oracle/__init__.py).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Every fixture is produced by calling the reference's own functions (`dropout_uncertainty`,
`select_samples`, `MyTransformer.forward`, `MLP.forward`, tokenizers, ...) with explicit seeds; the
dropout masks the reference drew are recorded through oracle/ref_shim.py and stored bit-packed so the
CUDA path can be driven with exactly the same masks (mask-injection mode).
"""

from __future__ import annotations

import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import restatement as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MODES = ("max", "margin", "entropy")


def sd_to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items() if k != "pos_enc.pe"}


def capture_last_logp(model, store):
    return model.register_forward_hook(lambda _m, _i, out: store.append(out[-1].detach().clone()))


def main_transformer(ref):
    """C2-style fixture: notebook-default BayesFormer, trained 40 epochs, B=32, T=20."""
    al, tr, pre, utils = ref.active_learning, ref.transformer, ref.preprocessing, ref.utils
    tok = ref.tokenizer.ReversedPairingTokenizer()
    random.seed(42)
    torch.manual_seed(42)
    train_set = pre.create_dataset(500, 3)
    valid_set = pre.create_dataset(100, 3)
    pool = pre.create_dataset(32, 3)
    model = tr.MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True)
    assert sum(p.numel() for p in model.parameters()) == 50178  # active_learning.ipynb cells 12-13
    utils.train(model, train_set, valid_set, 40, 20, 1e-3, tok.ntokens, tok, "cpu")
    model.eval()
    prompts, answers, P, A, _, _ = pre.get_batch(pool, tok, 0, len(pool))
    N, T, B, d, L = A + 1, 20, prompts.shape[1], 64, 2
    out = {"prompts": prompts.numpy(), "answers": answers.numpy(), "P": P, "N": N, "T": T, "p": 0.3}
    out.update({"sd." + k: v for k, v in sd_to_np(model.state_dict()).items()})

    # --- dropout scheme: record the reference's masks once, replay them for the other modes ---
    with torch.no_grad():
        log, logps = [], []
        hook = capture_last_logp(model, logps)
        torch.manual_seed(1234)
        with ref_shim.record_dropout(log):
            scores_entropy = al.dropout_uncertainty(model, prompts, N, "cpu", "entropy", T)
        hook.remove()
        assert len(log) == T * N * (1 + L)
        # order: for draw, for step: [E, LO_0, LO_1]  
        live = np.zeros((N, 1 + L, B * T, P + N - 1, 2), dtype=np.uint32)
        it = iter(log)
        for draw in range(T):
            for step in range(N):
                for site in range(1 + L):
                    keep = next(it).numpy()  # (P+step, B, d)
                    live[step, site, draw::T, : P + step, :] = R.pack_bits(np.transpose(keep, (1, 0, 2)))
        out["dropout.keep_bits"] = live  # [step][E, LO_0, LO_1][s = item*T + draw][pos][2]
        lp = torch.stack(logps).view(T, N, B, -1)  # (draw, step, item, V)
        out["dropout.logp_first_draws"] = lp[:6].numpy()
        out["dropout.tokens"] = lp.argmax(-1).numpy().astype(np.int16)  # (T, N, B) greedy continuation per draw
        out["dropout.scores.entropy"] = scores_entropy.numpy()
        for mode in ("max", "margin"):
            with ref_shim.replay_dropout(log):
                out[f"dropout.scores.{mode}"] = al.dropout_uncertainty(model, prompts, N, "cpu", mode, T).numpy()
        # select_samples through the reference's own entry point (list-of-pairs API), k = 8
        for mode in MODES:
            with ref_shim.replay_dropout(log):
                out[f"dropout.select.{mode}"] = np.array(
                    al.select_samples(model, tok, pool, "cpu", nb_samples=8, compute="dropout", mode=mode))

        # --- greedy scheme on the same weights without bayes (deterministic) ---
        model.bayes = False
        logps = []
        hook = capture_last_logp(model, logps)
        out["greedy.scores.entropy"] = al.greedy_uncertainty(model, prompts, N, "cpu", "entropy").numpy()
        hook.remove()
        out["greedy.logp"] = torch.stack(logps).numpy()  # (N, B, V)
        for mode in ("max", "margin"):
            out[f"greedy.scores.{mode}"] = al.greedy_uncertainty(model, prompts, N, "cpu", mode).numpy()

        # --- sampling scheme: record the multinomial tokens, replay them for the other modes ---
        drawn = []
        real_multinomial = torch.multinomial

        def rec(probs, num_samples=1, **kw):
            t = real_multinomial(probs, num_samples, **kw)
            drawn.append(t.view(-1).clone())
            return t

        torch.manual_seed(99)
        torch.multinomial = rec
        try:
            out["sampling.scores.entropy"] = al.sampling_uncertainty(model, prompts, N, "cpu", "entropy", 0.8, T).numpy()
        finally:
            torch.multinomial = real_multinomial
        forced = torch.stack(drawn).view(T, N, B)  # (draw, step, item)
        out["sampling.tokens"] = forced.numpy().astype(np.int16)
        for mode in ("max", "margin"):
            it2 = iter(drawn)
            torch.multinomial = lambda probs, num_samples=1, **kw: next(it2).view(-1, 1)
            try:
                out[f"sampling.scores.{mode}"] = al.sampling_uncertainty(model, prompts, N, "cpu", mode, 0.8, T).numpy()
            finally:
                torch.multinomial = real_multinomial
        model.bayes = True
    np.savez_compressed(os.path.join(OUT, "c2_transformer.npz"), **out)
    print("c2_transformer.npz", {k: getattr(v, "shape", v) for k, v in out.items() if not k.startswith("sd.")})


def latent_sites(ref):
    """Block-level ("latent") dropout sites switched on by hand: one forward pass."""
    tr = ref.transformer
    torch.manual_seed(5)
    V, d, H, F, L, n_pos, B, p = 20, 32, 4, 16, 2, 7, 3, 0.25
    model = tr.MyTransformer(V, d, H, F, L, bayes_dropout=p, bayes=True).eval()
    for layer in model.layers:
        layer.bayes, layer.bayes_dropout = True, p
        layer.ff.bayes_dropout = p
        layer.self_attn.bayes, layer.self_attn.bayes_dropout = True, p
    src = torch.randint(0, V, (n_pos, B))
    log = []
    with torch.no_grad(), ref_shim.record_dropout(log):
        logp = model(src)
    # order: E; per layer: HO_0..HO_{H-1} (n_pos,B,dh), attn-resid, FFN-in, FFN-resid, LO
    assert len(log) == 1 + L * (H + 4)
    packed = R.PackedMasks(1, L, B, n_pos, d)
    it = iter(log)
    packed.set(0, R.SITE_EMBED, next(it).numpy())
    for i in range(L):
        ho = np.concatenate([next(it).numpy() for _ in range(H)], axis=-1)
        packed.set(0, R.site_head_out(i), ho)
        packed.set(0, R.site_attn_resid(i), next(it).numpy())
        packed.set(0, R.site_ffn_in(i), next(it).numpy())
        packed.set(0, R.site_ffn_resid(i), next(it).numpy())
        packed.set(0, R.site_layer_out(i), next(it).numpy())
    out = {"src": src.numpy(), "logp": logp.numpy(), "bits": packed.bits, "p": p,
           "dims": np.array([V, d, H, F, L])}
    out.update({"sd." + k: v for k, v in sd_to_np(model.state_dict()).items()})
    np.savez_compressed(os.path.join(OUT, "latent_sites.npz"), **out)
    print("latent_sites.npz", logp.shape, len(log))


def head_in_sites(ref):
    """Head.bayes (reference transformer.py:50-53): every head built with bayes=True draws three independent input
    masks; together with MultiHeadAttention.bayes (head outputs) and the default sites.  One forward pass."""
    tr = ref.transformer
    torch.manual_seed(9)
    V, d, H, F, L, n_pos, B, p = 20, 32, 4, 16, 2, 6, 3, 0.25
    model = tr.MyTransformer(V, d, H, F, L, bayes_dropout=p, bayes=True).eval()
    for layer in model.layers:
        layer.self_attn.bayes, layer.self_attn.bayes_dropout = True, p
        for head in layer.self_attn.heads:
            head.bayes, head.bayes_dropout = True, p
    src = torch.randint(0, V, (n_pos, B))
    log = []
    with torch.no_grad(), ref_shim.record_dropout(log):
        logp = model(src)
    # order (verified by instrumenting F.dropout): E; per layer: per head Q, K, V inputs (n_pos,B,d) then that
    # head's output (n_pos,B,dh); LO
    assert len(log) == 1 + L * (4 * H + 1)
    packed = R.PackedMasks(1, L, B, n_pos, d, H)
    it = iter(log)
    packed.set(0, R.SITE_EMBED, next(it).numpy())
    for i in range(L):
        ho = []
        for h in range(H):
            for c in range(3):
                m = next(it).numpy()
                assert m.shape == (n_pos, B, d)
                packed.set(0, R.site_head_in(i, h, c, L, H), m)
            ho.append(next(it).numpy())
        packed.set(0, R.site_head_out(i), np.concatenate(ho, axis=-1))
        packed.set(0, R.site_layer_out(i), next(it).numpy())
    out = {"src": src.numpy(), "logp": logp.numpy(), "bits": packed.bits, "p": p, "dims": np.array([V, d, H, F, L])}
    out.update({"sd." + k: v for k, v in sd_to_np(model.state_dict()).items()})
    np.savez_compressed(os.path.join(OUT, "head_in_sites.npz"), **out)
    print("head_in_sites.npz", logp.shape, len(log))


def mid_model(ref):
    """dh=32, len=40 (two keys per lane in attention), L=3, p=0.2 — plain forward with recorded masks."""
    tr = ref.transformer
    torch.manual_seed(7)
    V, d, H, F, L, n_pos, B, p = 50, 64, 2, 128, 3, 40, 4, 0.2
    model = tr.MyTransformer(V, d, H, F, L, bayes_dropout=p, bayes=True).eval()
    src = torch.randint(0, V, (n_pos, B))
    log = []
    with torch.no_grad(), ref_shim.record_dropout(log):
        logp = model(src)
    packed = R.PackedMasks(1, L, B, n_pos, d)
    packed.set(0, R.SITE_EMBED, log[0].numpy())
    for i in range(L):
        packed.set(0, R.site_layer_out(i), log[1 + i].numpy())
    out = {"src": src.numpy(), "logp": logp.numpy(), "bits": packed.bits, "p": p, "dims": np.array([V, d, H, F, L])}
    out.update({"sd." + k: v for k, v in sd_to_np(model.state_dict()).items()})
    np.savez_compressed(os.path.join(OUT, "mid_model.npz"), **out)
    print("mid_model.npz", logp.shape)


def scaled_c4(ref):
    """C4 shape (d=512, H=8, F=2048, L=6, len=256): weights/masks are regenerated from the seeds on the
    GPU box (same image, same torch) — only inputs and the reference's outputs are stored."""
    tr = ref.transformer
    torch.manual_seed(11)
    V, d, H, F, L, n_pos, B, p = 114, 512, 8, 2048, 6, 256, 2, 0.3
    model = tr.MyTransformer(V, d, H, F, L, bayes_dropout=p, bayes=True).eval()
    assert sum(q.numel() for q in model.parameters()) == 19018866
    src = torch.randint(0, V, (n_pos, B))
    torch.manual_seed(12)
    with torch.no_grad():
        logp = model(src)
    out = {"src": src.numpy(), "logp_last": logp[-1].numpy(), "logp_rows": logp[::37].numpy(),
           "weight_seed": 11, "mask_seed": 12, "p": p, "dims": np.array([V, d, H, F, L]),
           "probe": model.layers[5].ff.net[2].weight.detach().numpy()[:4, :8]}
    np.savez_compressed(os.path.join(OUT, "scaled_c4.npz"), **out)
    print("scaled_c4.npz", logp.shape)


def char_level_c1(ref):
    """C1 shape (SURVEY 8d / BASELINE configs[0]): char-level model MyTransformer(67, 64, 8, 8, 2) over windows of the
    reference's data/shakespeare.txt (32-character prompts, 4-character answers: sequences of 32..36 tokens), a few
    epochs of the reference's own utils.train, then dropout_uncertainty with its masks recorded (B = 8, T = 6)."""
    al, tr, pre, utils = ref.active_learning, ref.transformer, ref.preprocessing, ref.utils
    constants = ref.constants
    text = open(os.path.join(ref_shim.REFERENCE_ROOT, "data", "shakespeare.txt")).read()
    chars = sorted(set(text))
    assert len(chars) == 65

    class CharTok(ref.tokenizer.Tokenizer):   # 65 characters + PAD + EOS, the interface preprocessing.pad / get_batch use
        def __init__(self):
            self.vocab = chars + [constants.PAD_TOKEN, constants.EOS_TOKEN]
            self.token_to_id = {c: i for i, c in enumerate(self.vocab)}
            self.ntokens = len(self.vocab)

        def encode(self, t):
            return [self.token_to_id[c] for c in t]

        def decode(self, ids):
            return "".join(self.vocab[i] for i in ids)

    tok = CharTok()
    windows = [(text[36 * i: 36 * i + 32], text[36 * i + 32: 36 * i + 36]) for i in range(400)]
    random.seed(42)
    torch.manual_seed(42)
    model = tr.MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True)
    utils.train(model, windows[8:328], windows[328:392], 6, 32, 2e-3, tok.ntokens, tok, "cpu")
    model.eval()
    pool = windows[:8]
    prompts, answers, P, A, _, _ = pre.get_batch(pool, tok, 0, len(pool))
    N, T, B, L = A + 1, 6, prompts.shape[1], 2
    assert (P, N) == (32, 5)
    out = {"prompts": prompts.numpy(), "P": P, "N": N, "T": T, "p": 0.3, "chars": np.array([ord(c) for c in chars])}
    out.update({"sd." + k: v for k, v in sd_to_np(model.state_dict()).items()})
    with torch.no_grad():
        log, logps = [], []
        hook = capture_last_logp(model, logps)
        torch.manual_seed(4321)
        with ref_shim.record_dropout(log):
            scores = al.dropout_uncertainty(model, prompts, N, "cpu", "entropy", T)
        hook.remove()
        assert len(log) == T * N * (1 + L)
        live = np.zeros((N, 1 + L, B * T, P + N - 1, 2), dtype=np.uint32)
        it = iter(log)
        for draw in range(T):
            for step in range(N):
                for site in range(1 + L):
                    keep = next(it).numpy()
                    live[step, site, draw::T, : P + step, :] = R.pack_bits(np.transpose(keep, (1, 0, 2)))
        out["dropout.keep_bits"] = live
        lp = torch.stack(logps).view(T, N, B, -1)
        out["dropout.logp"] = lp.numpy()
        out["dropout.tokens"] = lp.argmax(-1).numpy().astype(np.int16)
        out["dropout.scores.entropy"] = scores.numpy()
        for mode in ("max", "margin"):
            with ref_shim.replay_dropout(log):
                out[f"dropout.scores.{mode}"] = al.dropout_uncertainty(model, prompts, N, "cpu", mode, T).numpy()
    np.savez_compressed(os.path.join(OUT, "c1_char_level.npz"), **out)
    print("c1_char_level.npz", {k: getattr(v, "shape", v) for k, v in out.items() if not k.startswith("sd.")})


def classifiers(ref):
    bc, pc = ref.bayesian_classification, ref.preprocessing_classif
    X, _ = pc.get_data(256)
    T = 20
    torch.manual_seed(3)
    mlp = bc.MLP(2, 50, 1, dropout=True, dropout_rate=0.5).eval()
    log, outs = [], []
    with torch.no_grad(), ref_shim.record_dropout(log):
        for _ in range(T):  # visualization.py:110-112
            outs.append(torch.sigmoid(mlp(X)))
    probs = torch.stack(outs)  # (T, B, 1)
    keep = torch.stack(log).numpy()  # (T, B, 50)
    out = {"x": X.numpy(), "mlp.keep_bits": R.pack_bits(keep), "mlp.probs": probs.squeeze(-1).numpy(),
           "mlp.mean": probs.mean(0).squeeze().numpy(), "mlp.p": 0.5}
    out.update({"mlp.sd." + k: v.detach().numpy() for k, v in mlp.state_dict().items()})

    torch.manual_seed(4)
    vm = bc.VariationalMLP(2, 50, 1, prior_var=10.0).eval()
    with torch.no_grad():  # parameters are zero-initialised in the reference; give them signal
        for q in vm.parameters():
            q.copy_(torch.randn_like(q) * 0.7)
    eps_log = []
    real_randn_like = torch.randn_like

    def rec(t, **kw):
        e = real_randn_like(t, **kw)
        eps_log.append(e.reshape(-1).clone())
        return e

    torch.randn_like = rec
    outs = []
    try:
        with torch.no_grad():
            for _ in range(T):
                outs.append(torch.sigmoid(vm(X)))
    finally:
        torch.randn_like = real_randn_like
    probs = torch.stack(outs)
    eps = torch.stack([torch.cat(eps_log[2 * t: 2 * t + 2]) for t in range(T)])  # hidden layer first, then output
    out.update({"vmlp.eps": eps.numpy(), "vmlp.probs": probs.squeeze(-1).numpy(), "vmlp.mean": probs.mean(0).squeeze().numpy()})
    out.update({"vmlp.sd." + k: v.detach().numpy() for k, v in vm.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "classifiers.npz"), **out)
    print("classifiers.npz", probs.shape)


def uncertainty_and_tokenizer(ref):
    al, pre = ref.active_learning, ref.preprocessing
    torch.manual_seed(21)
    probs = torch.softmax(torch.randn(64, 114) * 2.0, dim=-1)
    probs2 = torch.softmax(torch.randn(33, 2), dim=-1)
    out = {"probs": probs.numpy(), "probs2": probs2.numpy()}
    for mode in MODES:
        out[mode] = al.compute_uncertainty(probs, mode).numpy()
        out[mode + "2"] = al.compute_uncertainty(probs2, mode).numpy()
    np.savez_compressed(os.path.join(OUT, "uncertainty.npz"), **out)

    random.seed(8)
    rp, ch = ref.tokenizer.ReversedPairingTokenizer(), ref.tokenizer.CharacterLevelTokenizer()
    data = pre.create_dataset(120, 3) + pre.create_dataset(40, 2) + pre.create_dataset(40, 5)
    cases = []
    for prompt, answer in data:
        full = rp.encode(prompt) + rp.encode(answer)
        cases.append({"prompt": prompt, "answer": answer, "rp_prompt": rp.encode(prompt), "rp_answer": rp.encode(answer),
                      "rp_decode_full": rp.decode(full), "ch_prompt": ch.encode(prompt), "ch_answer": ch.encode(answer)})
    batch = data[:9]
    Xb, Yb, Pb, Ab, _, _ = pre.get_batch(batch, rp, 0, len(batch))
    Xc, Yc, Pc, Ac, _, _ = pre.get_batch(batch, ch, 0, len(batch))
    js = {"cases": cases, "batch": batch, "rp_batch": {"X": Xb.tolist(), "Y": Yb.tolist(), "P": Pb, "A": Ab},
          "ch_batch": {"X": Xc.tolist(), "Y": Yc.tolist(), "P": Pc, "A": Ac},
          "rp_vocab": rp.vocabulary.vocab, "ch_vocab": ch.vocab,
          "decode_padded": rp.decode([112, 112, 73, 111, 100, 101, 113, 112])}
    with open(os.path.join(OUT, "tokenizer.json"), "w") as f:
        json.dump(js, f)
    print("uncertainty.npz, tokenizer.json", len(cases))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load_reference()
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["transformer", "latent", "headin", "mid", "c4", "c1", "classifiers", "misc"]
    if "transformer" in which:
        main_transformer(ref)
    if "latent" in which:
        latent_sites(ref)
    if "headin" in which:
        head_in_sites(ref)
    if "mid" in which:
        mid_model(ref)
    if "c4" in which:
        scaled_c4(ref)
    if "c1" in which:
        char_level_c1(ref)
    if "classifiers" in which:
        classifiers(ref)
    if "misc" in which:
        uncertainty_and_tokenizer(ref)


if __name__ == "__main__":
    main()
