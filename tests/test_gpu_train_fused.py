"""GPU parity of the hand-written fused fine-tune step (bayesformer_b200/csrc/train_fused.cu, `bfb200_train_step`)
against the reference's training arithmetic — eager autograd through the nn.Module stack + torch.optim.AdamW, the loop
of src/libs/utils.py:68-78 — with the dropout sites switched off (masks from different generators cannot be compared
element-wise): losses and every updated parameter must agree to 1e-4 after several steps."""

import copy
import random

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _batches(tok, n_batches, batch, seed):
    from bayesformer_b200.libs import preprocessing

    random.seed(seed)
    data = preprocessing.create_dataset(n_batches * batch, 3)
    out = []
    for i in range(n_batches):
        prompts, answers, p_len, a_len, _, _ = preprocessing.get_batch(data, tok, i * batch, batch)
        out.append((torch.cat((prompts, answers), 0), p_len))
    return out


@pytest.mark.parametrize("F,L", [(8, 2), (16, 1)])
def test_fused_step_matches_eager_autograd_and_adamw(dev, F, L):
    from bayesformer_b200 import _lib, engine
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
    from bayesformer_b200.model.transformer import MyTransformer

    _lib.load()
    tok = ReversedPairingTokenizer()
    torch.manual_seed(3)
    eager = MyTransformer(tok.ntokens, 64, 8, F, L, dropout=0.0).to(dev).train()
    fused = copy.deepcopy(eager)
    lr = 2e-3
    opt = torch.optim.AdamW(eager.parameters(), lr=lr)
    crit = nn.CrossEntropyLoss()
    trainer = engine.FusedTrainer(fused, lr)
    assert engine.FusedTrainer.supports(fused)
    batches = _batches(tok, 6, 20, seed=5)
    for tokens, p_len in batches:
        tokens = tokens.to(dev)
        eager.zero_grad()
        with torch.enable_grad():
            out = eager._module_forward(tokens)
            loss = crit(out[p_len - 1:-1].reshape(-1, tok.ntokens), tokens[p_len:].reshape(-1))
            loss.backward()
        opt.step()
        trainer.loss.zero_()
        trainer.step(tokens, p_len, 0.0, 0.0, seed=1)
        assert abs(float(trainer.loss.item()) - float(loss.item())) <= 1e-4 * max(1.0, abs(float(loss.item())))
    for (name, a), (_, b) in zip(eager.named_parameters(), fused.named_parameters()):
        # AdamW divides by sqrt(v): an element whose gradient is ~0 moves by up to lr per step whatever the gradient's
        # magnitude, so last-bit differences of such gradients surface as ~1e-5 absolute (1 % of one step of lr = 2e-3);
        # everything else agrees to 1e-4 relative
        np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().cpu().numpy(), rtol=1e-4, atol=3e-5, err_msg=name)
    # the first step's gradients, element-wise (fresh copies, one step)
    e2, f2 = copy.deepcopy(eager), copy.deepcopy(eager)
    tokens, p_len = batches[0][0].to(dev), batches[0][1]
    with torch.enable_grad():
        out = e2._module_forward(tokens)
        crit(out[p_len - 1:-1].reshape(-1, tok.ntokens), tokens[p_len:].reshape(-1)).backward()
    tr2 = engine.FusedTrainer(f2, lr)
    tr2.step(tokens, p_len, 0.0, 0.0, seed=1)
    lib = _lib.load()
    off = int(lib.bfb200_train_flat_offset(tok.ntokens, F, L, 0, 2))     # layer 0 W_out
    np.testing.assert_allclose(tr2.grad[off:off + 64 * 64].view(64, 64).cpu().numpy(),
                               e2.layers[0].self_attn.W_out.weight.grad.cpu().numpy(), rtol=1e-3, atol=1e-7)
    off = int(lib.bfb200_train_flat_offset(tok.ntokens, F, L, 0, 0))     # embedding (scatter-add over repeated tokens)
    np.testing.assert_allclose(tr2.grad[off:off + tok.ntokens * 64].view(-1, 64).cpu().numpy(),
                               e2.embedding.weight.grad.cpu().numpy(), rtol=1e-3, atol=1e-7)


def test_train_uses_the_fused_step_and_learns(dev):
    """utils.train on a CUDA device runs `train_seq_kernel` (one launch per batch, one CTA per sequence) for the notebook model, with the
    training-mode dropout sites live; the loss falls, the scoring packs follow the weights, and the graphed autograd
    path (BFB200_FUSED_TRAIN=0) reaches a comparable loss from the same start."""
    import os

    from bayesformer_b200 import _lib, engine
    from bayesformer_b200.libs import preprocessing, utils
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
    from bayesformer_b200.model.transformer import MyTransformer

    tok = ReversedPairingTokenizer()
    random.seed(11)
    train_set, valid_set = preprocessing.create_dataset(200, 3), preprocessing.create_dataset(40, 3)
    torch.manual_seed(4)
    model = MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).to(dev)
    twin = copy.deepcopy(model)
    prompts = preprocessing.get_batch(valid_set, tok, 0, 40)[0].to(dev)
    model.eval()
    before = engine.mc_score_transformer(engine.scoring_pack(model, dev, 4, 5), prompts, 5, 4, "dropout", "entropy", seed=9)
    with engine.profile(dev) as prof:
        tr_loss, va_loss = utils.train(model, train_set, valid_set, 8, 20, 2e-3, tok.ntokens, tok, dev)
    assert prof.kernels["train_seq_kernel"]["launches"] == 8 * 10
    assert tr_loss[-1] < 0.8 * tr_loss[0] and np.isfinite(va_loss).all()
    model.eval()
    after = engine.mc_score_transformer(engine.scoring_pack(model, dev, 4, 5), prompts, 5, 4, "dropout", "entropy", seed=9)
    assert not torch.allclose(after.step_scores, before.step_scores)
    os.environ["BFB200_FUSED_TRAIN"] = "0"
    try:
        tr2, va2 = utils.train(twin, train_set, valid_set, 8, 20, 2e-3, tok.ntokens, tok, dev)
    finally:
        del os.environ["BFB200_FUSED_TRAIN"]
    assert abs(va_loss[-1] - va2[-1]) <= 0.25 * va2[-1]       # different dropout streams: same training, statistically


def test_fused_step_gradient_with_dropout_matches_finite_differences(dev):
    """With the dropout sites ON the masks cannot be compared with torch's, but they are a deterministic function of
    (seed, step, row, site, feature): the loss of one step is an ordinary function of the weights, and the gradient the
    kernel computes (masks re-drawn in its backward pass) must be that function's derivative.  For EVERY parameter
    tensor: central difference of the kernel's own loss (lr = 0 leaves the weights untouched) along the direction
    sign(gradient) against sum |gradient| — a directional derivative, so the ReLU kinks and the fp32 rounding of the
    loss stay second-order."""
    from bayesformer_b200 import engine
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
    from bayesformer_b200.model.transformer import MyTransformer

    tok = ReversedPairingTokenizer()
    torch.manual_seed(8)
    model = MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).to(dev)
    tokens, p_len = _batches(tok, 1, 12, seed=2)[0]
    tokens = tokens.to(dev)
    trainer = engine.FusedTrainer(model, 0.0)
    trainer.WEIGHT_DECAY = 0.0

    def loss_and_grad():
        trainer.step_count = 0          # same step id => same masks
        trainer.loss.zero_()
        trainer.step(tokens, p_len, 0.1, 0.3, seed=77)
        return float(trainer.loss.item()), trainer.grad.clone()

    base, grad = loss_and_grad()
    again, grad2 = loss_and_grad()
    assert base == again and torch.equal(grad, grad2)     # the step is bit-reproducible
    eps = 1e-3
    checked = 0
    for param, seg in zip(trainer._params, trainer._segment_view()):
        g = grad[seg.offset:seg.offset + seg.count].view_as(param)
        want = float(g.abs().sum())
        if want < 1e-3:
            continue                     # e.g. a tensor whose gradient is (almost) masked out: nothing to resolve
        d = torch.sign(g)
        with torch.no_grad():
            param.add_(d, alpha=eps)
            up, _ = loss_and_grad()
            param.add_(d, alpha=-2 * eps)
            down, _ = loss_and_grad()
            param.add_(d, alpha=eps)
        fd = (up - down) / (2 * eps)
        assert abs(fd - want) <= 3e-2 * want + 2e-3, (tuple(param.shape), seg.offset, fd, want)
        checked += 1
    assert checked >= 40                 # per-head projections, W_out, both LayerNorms, FFN, embedding, head


def test_fused_step_large_batch_and_longest_sequence(dev):
    """250 sequences of 16 tokens (the class limit) in one step = 250 CTAs and gradient slabs; against eager autograd."""
    from bayesformer_b200 import engine
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
    from bayesformer_b200.model.transformer import MyTransformer

    tok = ReversedPairingTokenizer()
    torch.manual_seed(13)
    eager = MyTransformer(tok.ntokens, 64, 8, 8, 2, dropout=0.0).to(dev).train()
    fused = copy.deepcopy(eager)
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(0, tok.ntokens, (16, 250), generator=g).to(dev)
    p_len = 9
    with torch.enable_grad():
        out = eager._module_forward(tokens)
        loss = nn.CrossEntropyLoss()(out[p_len - 1:-1].reshape(-1, tok.ntokens), tokens[p_len:].reshape(-1))
        loss.backward()
    trainer = engine.FusedTrainer(fused, 1e-3)
    trainer.step(tokens, p_len, 0.0, 0.0, seed=1)
    assert abs(float(trainer.loss.item()) - float(loss.item())) <= 1e-4 * abs(float(loss.item()))
    ref = torch.cat([p.grad.reshape(-1) for p in (eager.embedding.weight,)])
    np.testing.assert_allclose(trainer.grad[:ref.numel()].cpu().numpy(), ref.cpu().numpy(), rtol=2e-3, atol=2e-7)
    with pytest.raises(ValueError):
        trainer.step(torch.zeros((17, 4), dtype=torch.int64, device=dev), 8, 0.0, 0.0, seed=1)   # beyond the class: caller falls back
