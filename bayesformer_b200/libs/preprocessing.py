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


def get_batch(dataset: list[tuple[str, str]], tokenizer: Tokenizer, i: int, batch_size: int):
    """Encode `dataset[i : i + batch_size]` → `(X (P,B), Y (A+1,B), P, A, prompts, answers)`."""
    rows = [dataset[j] for j in range(i, i + batch_size)]
    prompts = [row[0] for row in rows]
    answers = [row[1] for row in rows]
    padded_prompts, prompt_length = pad(tokenizer, [tokenizer.encode(p) for p in prompts], "prompts")
    padded_answers, answers_length = pad(tokenizer, [tokenizer.encode(a) for a in answers], "answers")
    X = torch.tensor(padded_prompts, dtype=torch.int64).t().contiguous()
    Y = torch.tensor(padded_answers, dtype=torch.int64).t().contiguous()
    return X, Y, prompt_length, answers_length, prompts, answers
