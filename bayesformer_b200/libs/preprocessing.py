"""Addition-problem datasets and batch assembly.

Mirror of the reference's `src/libs/preprocessing.py` (`sample_datapoint` :9,
`create_dataset` :29, `pad` :43, `get_batch` :75).  `get_batch` keeps the exact
layout contract the hot path depends on: prompts `(P, B)` int64 **left** padded
with PAD, answers `(A+1, B)` = ids + EOS, right padded with PAD, where `P` and
`A` are maxima over the slice (:59).  The RNG consumption of `sample_datapoint`
(2·num_digits `random.randint(0, 9)` calls, first operand first) is preserved so
`random.seed(42)` reproduces the reference's datasets.
"""

from __future__ import annotations

import random

import torch

from bayesformer_b200.configs import constants
from bayesformer_b200.libs.tokenizer import Tokenizer


def sample_datapoint(num_digits: int = 3) -> tuple[str, str]:
    """One `("abc+def=", "sum")` pair; operands keep leading zeros, the sum does not."""
    lhs = "".join(str(random.randint(0, 9)) for _ in range(num_digits))
    rhs = "".join(str(random.randint(0, 9)) for _ in range(num_digits))
    return f"{lhs}+{rhs}=", str(int(lhs) + int(rhs))


def create_dataset(nb_samples: int, num_digits: int = 3) -> list[tuple[str, str]]:
    return [sample_datapoint(num_digits) for _ in range(nb_samples)]


def pad(tokenizer: Tokenizer, token_list: list[list[int]], type_list: str = "prompts"):
    """Pad id lists to a common length; returns `(padded, max_length)` (reference :43-72).

    "prompts": PAD on the left.  "answers": EOS appended, then PAD on the right, so
    rows have `max_length + 1` entries.  Any other `type_list` yields an empty list,
    as in the reference.
    """
    pad_id = tokenizer.token_to_id[constants.PAD_TOKEN]
    eos_id = tokenizer.token_to_id[constants.EOS_TOKEN]
    max_length = max(len(ids) for ids in token_list)
    if type_list == "prompts":
        out = [[pad_id] * (max_length - len(ids)) + ids for ids in token_list]
    elif type_list == "answers":
        out = [ids + [eos_id] + [pad_id] * (max_length - len(ids)) for ids in token_list]
    else:
        out = []
    return out, max_length


def _encode_addition_pool(prompts: list[str], answers: list[str], tokenizer) -> tuple | None:
    """Vectorised `ReversedPairingTokenizer` encoding of a whole pool of canonical `"ddd+ddd="` / `"ddd"` pairs
    (the reference's per-item Python loop tokenises 45 k items/s; a 1 M pool would take 22 s per
    acquisition round).  Returns `(X (P,B), Y (A+1,B), P, A)` bit-identical to the per-item path, or None when the
    strings are not all of one canonical width — the caller then falls back to the per-item encoder.

    Encoding restated from the reference (`src/libs/tokenizer.py:177-247`): both operands lose their leading
    zeros, the shorter is left-padded with zeros to the longer, digit pairs are emitted units first as the token
    `10*a + b`, then `'='`; an answer is its digits reversed (ids 100 + digit); prompts are LEFT padded with PAD,
    answers get EOS and are RIGHT padded (`src/libs/preprocessing.py:59-72`)."""
    import numpy as np

    if not prompts:
        return None
    width = len(prompts[0])
    n = (width - 2) // 2
    if n < 1 or width != 2 * n + 2:
        return None
    try:
        raw = np.frombuffer("".join(prompts).encode("ascii"), dtype=np.uint8)
    except UnicodeEncodeError:
        return None
    if raw.size != len(prompts) * width:
        return None
    raw = raw.reshape(len(prompts), width)
    lhs, plus, rhs, eq = raw[:, :n], raw[:, n], raw[:, n + 1:2 * n + 1], raw[:, 2 * n + 1]
    digits_ok = ((lhs >= 48) & (lhs <= 57)).all() and ((rhs >= 48) & (rhs <= 57)).all()
    if not (digits_ok and (plus == 43).all() and (eq == 61).all()):
        return None
    t2i = tokenizer.token_to_id
    pad_id, eos_id, eq_id = t2i[constants.PAD_TOKEN], t2i[constants.EOS_TOKEN], t2i["="]
    if any(t2i[f"{k:02d}"] != k for k in (0, 7, 42, 99)) or any(t2i[str(k)] != 100 + k for k in range(10)):
        return None   # not the reference vocabulary layout
    a, b = (lhs - 48).astype(np.int64), (rhs - 48).astype(np.int64)          # most significant digit first
    sig = lambda d: np.maximum(n - (np.cumsum(d, axis=1) == 0).sum(axis=1), 1)   # length without leading zeros
    pairs = np.maximum(sig(a), sig(b))
    tok = a[:, ::-1] * 10 + b[:, ::-1]                                        # units first
    P = int(pairs.max()) + 1
    X = np.full((len(prompts), P), pad_id, dtype=np.int64)
    for w in range(1, n + 1):
        rows = np.nonzero(pairs == w)[0]
        if rows.size:
            X[rows, P - 1 - w:P - 1] = tok[rows, :w]
    X[:, P - 1] = eq_id
    # answers: plain digit strings
    lens = np.fromiter(map(len, answers), dtype=np.int64, count=len(answers))
    A = int(lens.max())
    if int(lens.min()) < 1:
        return None
    try:
        flat = np.frombuffer("".join(s.rjust(A, "x") for s in answers).encode("ascii"), dtype=np.uint8)
    except UnicodeEncodeError:
        return None
    if flat.size != len(answers) * A:
        return None
    flat = flat.reshape(len(answers), A)                                      # right-aligned, 'x' = 120 on the left
    isdig = (flat >= 48) & (flat <= 57)
    if not (isdig | (flat == 120)).all() or not (isdig.sum(axis=1) == lens).all():
        return None
    Y = np.full((len(answers), A + 1), pad_id, dtype=np.int64)
    rev = flat[:, ::-1]                                                       # units first, 'x' padding on the right
    Y[:, :A] = np.where(rev == 120, pad_id, rev.astype(np.int64) - 48 + 100)
    Y[np.arange(len(answers)), lens] = eos_id
    return (torch.from_numpy(np.ascontiguousarray(X.T)), torch.from_numpy(np.ascontiguousarray(Y.T)), P - 0, A)


def _encode_char_pool(prompts: list[str], answers: list[str], tokenizer) -> tuple | None:
    """Vectorised encoding of a whole pool under a character-level tokenizer (`CharacterLevelTokenizer`, reference
    `src/libs/tokenizer.py:28-62`, or the char-vocabulary tokenizer of the Shakespeare-window pool): one id per
    character, characters outside the vocabulary dropped (the reference's `clean`), prompts LEFT padded with PAD,
    answers + EOS RIGHT padded (`src/libs/preprocessing.py:59-72`).  Ragged and empty strings are fine.  Returns
    `(X (P,B), Y (A+1,B), P, A)` bit-identical to the per-item path, or None when the vocabulary is not one character
    per token."""
    import numpy as np

    t2i = tokenizer.token_to_id
    single = {ord(k): v for k, v in t2i.items() if len(k) == 1}
    if not prompts or len(single) != len(t2i) - 2 or max(single) > 0xFFFF:
        return None
    pad_id, eos_id = t2i[constants.PAD_TOKEN], t2i[constants.EOS_TOKEN]
    table = np.full(0x10000, -1, dtype=np.int64)
    table[np.fromiter(single.keys(), dtype=np.int64)] = np.fromiter(single.values(), dtype=np.int64)
    B = len(prompts)

    def encode(strings):
        lens = np.fromiter(map(len, strings), dtype=np.int64, count=B)
        flat = np.frombuffer("".join(strings).encode("utf-32-le", "surrogatepass"), dtype=np.uint32)
        if flat.size != int(lens.sum()):
            return None
        ids = np.where(flat < 0x10000, table[np.minimum(flat, 0xFFFF)], -1)
        keep = ids >= 0
        rows = np.repeat(np.arange(B), lens)[keep]
        ids = ids[keep]
        counts = np.bincount(rows, minlength=B)
        pos = np.arange(ids.size) - (np.cumsum(counts) - counts)[rows]
        return rows, ids, counts, pos

    ep, ea = encode(prompts), encode(answers)
    if ep is None or ea is None:
        return None
    rows, ids, counts, pos = ep
    P = int(counts.max())
    X = np.full((B, P), pad_id, dtype=np.int64)
    X[rows, P - counts[rows] + pos] = ids
    rows, ids, counts, pos = ea
    A = int(counts.max())
    Y = np.full((B, A + 1), pad_id, dtype=np.int64)
    Y[rows, pos] = ids
    Y[np.arange(B), counts] = eos_id
    return (torch.from_numpy(np.ascontiguousarray(X.T)), torch.from_numpy(np.ascontiguousarray(Y.T)), P, A)


def get_batch(dataset: list[tuple[str, str]], tokenizer: Tokenizer, i: int, batch_size: int):
    """Encode `dataset[i : i + batch_size]` → `(X (P,B), Y (A+1,B), P, A, prompts, answers)`.

    Pools of canonical addition strings under the `ReversedPairingTokenizer` are encoded by the vectorised
    `_encode_addition_pool`, pools under a character-level tokenizer by `_encode_char_pool` (identical output,
    ~25-100x faster); anything else takes the reference's per-item route."""
    rows = [dataset[j] for j in range(i, i + batch_size)]
    prompts = [row[0] for row in rows]
    answers = [row[1] for row in rows]
    if batch_size >= 64 and type(tokenizer).__name__ == "ReversedPairingTokenizer":
        fast = _encode_addition_pool(prompts, answers, tokenizer)
        if fast is not None:
            X, Y, prompt_length, answers_length = fast
            return X, Y, prompt_length, answers_length, prompts, answers
    if batch_size >= 64 and type(tokenizer).__name__ in ("CharacterLevelTokenizer", "CharVocabTokenizer"):
        fast = _encode_char_pool(prompts, answers, tokenizer)
        if fast is not None:
            X, Y, prompt_length, answers_length = fast
            return X, Y, prompt_length, answers_length, prompts, answers
    padded_prompts, prompt_length = pad(tokenizer, [tokenizer.encode(p) for p in prompts], "prompts")
    padded_answers, answers_length = pad(tokenizer, [tokenizer.encode(a) for a in answers], "answers")
    X = torch.tensor(padded_prompts, dtype=torch.int64).t().contiguous()
    Y = torch.tensor(padded_answers, dtype=torch.int64).t().contiguous()
    return X, Y, prompt_length, answers_length, prompts, answers
