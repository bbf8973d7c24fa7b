"""CPU-side tests: C-ABI library loads and exports every declared symbol, host mirrors of the reference's
tokenizer / batch layout reproduce its known answers, bench inputs match the tokenizer, and the product
refuses to run the scoring path without CUDA (no CPU fallback)."""

import json
import os
import random
import re

import numpy as np
import pytest
import torch

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from bayesformer_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "bayesformer_b200.h")).read()
    declared = set(re.findall(r"\b(bfb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.bfb200_abi_version() == _lib.ABI_VERSION
    # argument validation happens before any device work
    assert lib.bfb200_compute_uncertainty(None, 4, 3, 7, None, None) == _lib.ERR_INVALID_ARG
    assert b"Unknown mode" in lib.bfb200_last_error()
    assert lib.bfb200_topk_workspace_bytes(1_000_000, 200) > 0


def test_struct_layouts_match_header_sizes():
    import ctypes as C

    from bayesformer_b200 import _lib

    # 6 int32 + float + uint32 (32 B) + 14 pointers (112 B) + float + int32 (8 B) + 1 pointer
    assert C.sizeof(_lib.Transformer) == 160
    assert C.sizeof(_lib.Masks) == 24
    assert C.sizeof(_lib.McArgs) == 8 + 8 + 4 * 5 + 4 + 8 + 24 + 8 * 5


def test_no_cpu_path():
    from bayesformer_b200.libs import active_learning as al
    from bayesformer_b200.model.transformer import MyTransformer

    m = MyTransformer(114, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).eval()
    prompts = torch.zeros(4, 3, dtype=torch.int64)
    for fn in (al.greedy_uncertainty, al.sampling_uncertainty, al.dropout_uncertainty):
        with pytest.raises(RuntimeError, match="no CPU"):
            fn(m, prompts, 5, "cpu")
    with pytest.raises(RuntimeError, match="no CPU"):
        al.compute_uncertainty(torch.rand(3, 4), "max")
    with pytest.raises(ValueError):
        al.compute_uncertainty(torch.rand(3, 4), "nope")
    with pytest.raises(ValueError):
        al.select_samples_from_tokens(m, prompts, 5, "cpu", 2, compute="nope")
    assert al.create_new_train_set([("a", "b")], [("c", "d"), ("e", "f")], [1]) == [("a", "b"), ("e", "f")]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from bayesformer_b200 import _lib

    monkeypatch.setenv("BAYESFORMER_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_module_contract_matches_reference():
    """state_dict names / parameter count of active_learning.ipynb cells 12-13 (50,178) and autograd."""
    import copy

    from bayesformer_b200.model.transformer import MyTransformer

    m = MyTransformer(114, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True)
    assert sum(p.numel() for p in m.parameters()) == 50178
    z = H.load("c2_transformer.npz")
    names = {k[3:] for k in z.files if k.startswith("sd.")} | {"pos_enc.pe"}
    assert set(m.state_dict().keys()) == names
    m2 = copy.deepcopy(m)
    out = m2(torch.randint(0, 114, (5, 3)))
    out.sum().backward()
    assert m2.embedding.weight.grad is not None and out.shape == (5, 3, 114)


def test_module_forward_matches_reference_outputs():
    """The autograd nn.Module stack (training path) reproduces the reference's stored greedy log-probs."""
    z = H.load("c2_transformer.npz")
    model = H.build_model(H.state_dict_from(z), None, False)
    with torch.no_grad():
        out = model(torch.from_numpy(z["prompts"]))
    np.testing.assert_allclose(out[-1].numpy(), z["greedy.logp"][0], rtol=2e-5, atol=2e-5)


def test_tokenizer_and_get_batch_known_answers():
    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import CharacterLevelTokenizer, ReversedPairingTokenizer

    rp, ch = ReversedPairingTokenizer(), CharacterLevelTokenizer()
    assert rp.ntokens == 114 and ch.ntokens == 14
    assert rp.encode("21+352=") == [12, 25, 3, 111]          # SURVEY §4 known answers
    assert rp.encode("440+420=") == [0, 42, 44, 111]
    assert rp.encode("007+003=") == [73, 111]
    assert rp.encode("1560") == [100, 106, 105, 101]
    X, Y, P, A, _, _ = preprocessing.get_batch([("440+420=", "860"), ("007+003=", "10"), ("937+623=", "1560")], rp, 0, 3)
    assert X.t().tolist() == [[0, 42, 44, 111], [112, 112, 73, 111], [73, 32, 96, 111]]
    assert Y.t().tolist() == [[100, 106, 108, 113, 112], [100, 101, 113, 112, 112], [100, 106, 105, 101, 113]]
    assert (P, A) == (4, 4) and X.dtype == torch.int64
    js = json.load(open(os.path.join(H.GOLDEN, "tokenizer.json")))
    for c in js["cases"]:
        assert rp.encode(c["prompt"]) == c["rp_prompt"] and rp.encode(c["answer"]) == c["rp_answer"]
        assert rp.decode(c["rp_prompt"] + c["rp_answer"]) == c["rp_decode_full"]
        assert ch.encode(c["prompt"]) == c["ch_prompt"] and ch.encode(c["answer"]) == c["ch_answer"]
    batch = [tuple(b) for b in js["batch"]]
    Xb, Yb, Pb, Ab, _, _ = preprocessing.get_batch(batch, rp, 0, len(batch))
    assert Xb.tolist() == js["rp_batch"]["X"] and Yb.tolist() == js["rp_batch"]["Y"]
    assert (Pb, Ab) == (js["rp_batch"]["P"], js["rp_batch"]["A"])
    assert rp.vocabulary.vocab == js["rp_vocab"] and ch.vocab == js["ch_vocab"]
    assert rp.decode([112, 112, 73, 111, 100, 101, 113, 112]) == js["decode_padded"]


def test_bench_synthetic_prompts_equal_tokenizer_output():
    import bench
    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    n = 3000
    got = bench.synthetic_addition_prompts(n, seed=7)
    rng = np.random.RandomState(7)
    digits = rng.randint(0, 10, size=(n, 2, 3))
    pool = [("".join(map(str, d[0])) + "+" + "".join(map(str, d[1])) + "=", "0") for d in digits]
    X, _, P, _, _, _ = preprocessing.get_batch(pool, ReversedPairingTokenizer(), 0, n)
    assert P == 4 and np.array_equal(got, X.numpy())
    assert bench.flops_per_pool_sample(4, 5) == pytest.approx(44.36e6, rel=1e-3)


def test_shard_bounds_cover_pool():
    from bayesformer_b200 import sharding

    for n, g in ((1_000_000, 8), (10, 3), (7, 8), (0, 2)):
        spans = [sharding.shard_bounds(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)
    assert sharding.local_candidates(200, 50) == 50 and sharding.local_candidates(200, 5000) == 200


def test_vectorised_pool_tokenisation_is_bit_identical():
    """whole-pool encoding == the reference's per-item encode + pad, including short operands,
    leading zeros, 2- and 5-digit problems and the fallback for irregular strings."""
    import time

    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    tok = ReversedPairingTokenizer()

    def slow(data):
        P = [tok.encode(p) for p, _ in data]
        A = [tok.encode(a) for _, a in data]
        Xp, pl = preprocessing.pad(tok, P, "prompts")
        Yp, al = preprocessing.pad(tok, A, "answers")
        return torch.tensor(Xp).t(), torch.tensor(Yp).t(), pl, al

    for digits, n in ((3, 5000), (2, 300), (5, 300), (1, 100)):
        random.seed(digits)
        data = preprocessing.create_dataset(n, digits)
        data[0] = ("0" * digits + "+" + "0" * digits + "=", "0")
        X, Y, P, A, _, _ = preprocessing.get_batch(data, tok, 0, n)
        Xs, Ys, Ps, As = slow(data)
        assert (P, A) == (Ps, As) and torch.equal(X, Xs) and torch.equal(Y, Ys)
    # all-short operands: P shrinks exactly as the per-item path's maximum does
    data = [("007+003=", "10"), ("000+009=", "9")] * 40
    X, Y, P, A, _, _ = preprocessing.get_batch(data, tok, 0, len(data))
    Xs, Ys, Ps, As = slow(data)
    assert (P, A) == (Ps, As) == (2, 2) and torch.equal(X, Xs) and torch.equal(Y, Ys)
    # irregular strings fall back to the per-item encoder
    mixed = [("12+345=", "357")] + [("440+420=", "860")] * 70
    X, Y, P, A, _, _ = preprocessing.get_batch(mixed, tok, 0, len(mixed))
    Xs, Ys, Ps, As = slow(mixed)
    assert (P, A) == (Ps, As) and torch.equal(X, Xs) and torch.equal(Y, Ys)
    random.seed(1)
    big = preprocessing.create_dataset(100_000, 3)
    t0 = time.perf_counter()
    preprocessing.get_batch(big, tok, 0, len(big))
    assert time.perf_counter() - t0 < 2.0   # the per-item path needs ~2.2 s for 100 k items


def test_repad_prompts_matches_whole_pool_padding():
    """A shard whose prompts are all short is padded up to the pool-wide prompt length, exactly as the reference's
    whole-pool get_batch would have padded it."""
    from bayesformer_b200 import sharding
    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    tok = ReversedPairingTokenizer()
    pool = [("007+003=", "10"), ("000+009=", "9")] * 40 + [("440+420=", "860")] * 80
    X, _, P, _, _, _ = preprocessing.get_batch(pool, tok, 0, len(pool))
    Xa, _, Pa, _, _, _ = preprocessing.get_batch(pool, tok, 0, 80)
    assert Pa < P
    assert torch.equal(sharding.repad_prompts(Xa, P, tok.token_to_id["<PAD>"] if "<PAD>" in tok.token_to_id else 112), X[:, :80])
    with pytest.raises(ValueError):
        sharding.repad_prompts(X, Pa, 112)


def test_vectorised_char_pool_tokenisation_is_bit_identical():
    """Character-level tokenizers (the reference's CharacterLevelTokenizer and the 67-symbol tokenizer of the
    Shakespeare-window pool): ragged and empty strings, characters outside the vocabulary (dropped, as `clean` does)."""
    import random
    import string

    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import CharacterLevelTokenizer, CharVocabTokenizer

    rnd = random.Random(7)
    alphabet = string.ascii_letters + " ,.;:!?'\n-&3$"
    for tok, chars in ((CharVocabTokenizer(alphabet), alphabet + "#@é"), (CharacterLevelTokenizer(), "0123456789+= x")):
        pool = [("".join(rnd.choice(chars) for _ in range(rnd.choice((0, 1, 5, 32, 33)))),
                 "".join(rnd.choice(chars) for _ in range(rnd.choice((0, 2, 4))))) for _ in range(300)]
        fast = preprocessing.get_batch(pool, tok, 0, len(pool))
        slow_p, p_len = preprocessing.pad(tok, [tok.encode(p) for p, _ in pool], "prompts")
        slow_a, a_len = preprocessing.pad(tok, [tok.encode(a) for _, a in pool], "answers")
        assert (fast[2], fast[3]) == (p_len, a_len)
        assert torch.equal(fast[0], torch.tensor(slow_p, dtype=torch.int64).t())
        assert torch.equal(fast[1], torch.tensor(slow_a, dtype=torch.int64).t())
        # small batches keep the per-item route and agree as well
        few = preprocessing.get_batch(pool, tok, 10, 20)
        sp, _ = preprocessing.pad(tok, [tok.encode(p) for p, _ in pool[10:30]], "prompts")
        assert torch.equal(few[0], torch.tensor(sp, dtype=torch.int64).t())


def test_weight_image_signature_tracks_versions_sites_and_explicit_marks():
    """The packed / 16-bit weight images are keyed on every parameter's (data_ptr, _version), the dropout-site
    configuration and an explicit generation counter (graph replays of an optimizer step update weights without
    touching `_version`: utils.train bumps the counter)."""
    from bayesformer_b200 import engine
    from bayesformer_b200.model.transformer import MyTransformer

    torch.manual_seed(0)
    m = MyTransformer(20, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True)
    dev = torch.device("cpu")
    s0 = engine._param_signature(m, dev)
    assert engine._param_signature(m, dev) == s0
    with torch.no_grad():
        m.linear_out.bias.add_(1.0)                     # in-place update through the dispatcher bumps _version
    s1 = engine._param_signature(m, dev)
    assert s1 != s0
    engine.mark_weights_changed(m)                       # an update behind autograd's back
    s2 = engine._param_signature(m, dev)
    assert s2 != s1
    for layer in m.layers:
        layer.ff.bayes_dropout = 0.3                     # a latent dropout site switched on (in every layer)
    assert engine._param_signature(m, dev) != s2


def test_train_step_layout_functions_are_consistent():
    """The pure layout functions of the fused fine-tune step (no GPU): the flat layout covers every parameter of the
    notebook model exactly once in the documented order, and shapes outside the step's class are reported by a zero
    scratch size (the caller then trains on another path) rather than by a launch failure."""
    from bayesformer_b200 import _lib
    from bayesformer_b200.model.transformer import MyTransformer

    lib = _lib.load()
    V, F, L = 114, 8, 2
    model = MyTransformer(V, 64, 8, F, L, bayes_dropout=0.3, bayes=True)
    n = int(lib.bfb200_train_state_floats(V, F, L))
    assert n == sum(p.numel() for p in model.parameters())
    off = lambda layer, which: int(lib.bfb200_train_flat_offset(V, F, L, layer, which))
    sizes = {0: V * 64, 1: 3 * 64 * 64, 2: 64 * 64, 3: 64, 4: 64, 5: F * 64, 6: F, 7: 64 * F, 8: 64, 9: 64, 10: 64,
             11: V * 64, 12: V}
    spans = [(off(0, 0), sizes[0])]
    for layer in range(L):
        spans += [(off(layer, w), sizes[w]) for w in range(1, 11)]
    spans += [(off(0, 11), sizes[11]), (off(0, 12), sizes[12])]
    spans.sort()
    assert spans[0][0] == 0 and all(a + s == b for (a, s), (b, _) in zip(spans, spans[1:])) and spans[-1][0] + spans[-1][1] == n
    assert all(o % 4 == 0 for o, s in spans if s >= 4)                 # float4 access in the kernel
    assert int(lib.bfb200_train_flat_offset(V, F, L, 0, 13)) == -1
    ok = int(lib.bfb200_train_scratch_floats(20, 9, 4, V, F, L))
    assert ok > 20 * n                                                  # one gradient slab per sequence
    assert int(lib.bfb200_train_scratch_floats(40, 9, 4, V, F, L)) > ok
    for bad in ((20, 17, 4, V, F, L),      # longer than the class's 16 tokens
                (20, 9, 9, V, F, L),       # no answer rows
                (0, 9, 4, V, F, L),
                (20, 9, 4, V, 6, L),       # d_ff not a multiple of 4
                (20, 9, 4, 400, F, L),     # vocabulary beyond the head stage
                (20, 16, 4, V, 64, 12)):   # activations of one sequence beyond a CTA's shared memory
        assert int(lib.bfb200_train_scratch_floats(*bad)) == 0, bad
