"""Tokenizers feeding the MC-scoring path.

Behavioural mirror of the reference's `src/libs/tokenizer.py` (Tokenizer ABC :9-18,
CharacterLevelTokenizer :28-62, ReversedPairingTokenizer :343-380) written from its
observable behaviour: same vocabularies and ids (14 and 114 symbols), same
cleaning rule, same "reversed digit pair" encoding of `abc+def=`.  Known answers
from the reference docstrings (:221-223, :289) are pinned in tests/test_host_logic.py
(test_tokenizer_and_get_batch_known_answers) and tests/golden/tokenizer.json.

`encode_pool` / `encode_batch_numpy` are the vectorised entry points used by the
pool-scoring path: they tokenise a whole pool once, instead of the
reference's per-item Python loop in `preprocessing.get_batch`.
"""

from __future__ import annotations

import abc

import regex as re

from bayesformer_b200.configs import constants


class Tokenizer(abc.ABC):
    """Interface every tokenizer of the repo implements (reference :9-18)."""

    @abc.abstractmethod
    def encode(self, text: str) -> list[int]:
        raise NotImplementedError

    @abc.abstractmethod
    def decode(self, token_list: list[int]) -> str:
        raise NotImplementedError


def _charset_filter(symbols: list[str]) -> str:
    # The reference builds one negated character class from the *characters* of
    # all symbols joined together (:42, :133), so the letters of "<PAD>"/"<EOS>"
    # survive cleaning.  Kept as is: it is observable through encode().
    return f"[^{re.escape(''.join(symbols))}]"


class CharacterLevelTokenizer(Tokenizer):
    """One id per character: digits, '+', '=', PAD, EOS (reference :28-62)."""

    def __init__(self):
        digits = [str(i) for i in range(10)]
        self.vocab = digits + ["+", "="] + [constants.PAD_TOKEN, constants.EOS_TOKEN]
        self.token_to_id = {tok: i for i, tok in enumerate(self.vocab)}
        self.id_to_token = dict(enumerate(self.vocab))
        self.ntokens = len(self.vocab)
        self.pattern = _charset_filter(self.vocab)

    def clean(self, text: str) -> str:
        return re.sub(self.pattern, "", text)

    def pre_tokenization(self, text: str) -> list[str]:
        return list(text)

    def encode(self, text: str) -> list[int]:
        return [self.token_to_id[ch] for ch in self.pre_tokenization(self.clean(text))]

    def decode(self, token_list: list[int]) -> str:
        return "".join(self.id_to_token[i] for i in token_list)


# ---------------------------------------------------------------------------
# Reversed digit-pair tokenizer
# ---------------------------------------------------------------------------


def remove_leading_zeros(s: str) -> str:
    """"007" -> "7", "000" -> "0" (reference :72-82)."""
    stripped = s.lstrip("0")
    return stripped if stripped else "0"


def pad_with_0(s: str, m: int) -> str:
    """Left-pad `s` with zeros up to length `m` (never truncates; reference :85-99)."""
    return s.rjust(m, "0") if len(s) < m else s


def pad_equally_0(s1: str, s2: str) -> tuple[str, str]:
    """Strip leading zeros of both operands then left-pad the shorter (reference :102-119)."""
    a, b = remove_leading_zeros(s1), remove_leading_zeros(s2)
    width = max(len(a), len(b))
    return pad_with_0(a, width), pad_with_0(b, width)


class Vocabulary:
    """100 digit pairs "00".."99", 10 digits, '+', '=', PAD, EOS → 114 ids (reference :122-174)."""

    def __init__(self, method: str = "tens_pairing"):
        self.vocab = self.get_vocab(method=method)
        self.token_to_id = {tok: i for i, tok in enumerate(self.vocab)}
        self.id_to_token = dict(enumerate(self.vocab))
        self.pattern = _charset_filter(self.vocab)
        self.ntokens = len(self.vocab)

    def get_vocab(self, method: str) -> list[str]:
        if method != "tens_pairing":
            raise NotImplementedError
        pairs = [f"{i:02d}" for i in range(100)]
        digits = [str(i) for i in range(10)]
        return pairs + digits + ["+", "="] + [constants.PAD_TOKEN, constants.EOS_TOKEN]

    def clean(self, text: str) -> str:
        return re.sub(self.pattern, "", text)


class Encoder:
    """`"21+352="` → ids of "12","25","03" (units first) + '=' ; `"860"` → digit ids reversed (reference :177-247)."""

    def __init__(self, vocabulary: Vocabulary):
        self.vocabulary = vocabulary

    def _encode_number(self, number_str: str) -> list[int]:
        t2i = self.vocabulary.token_to_id
        return [t2i[ch] for ch in reversed(number_str)]

    def _encode_addition(self, addition_str: str) -> list[int]:
        lhs, rhs = addition_str.split("+")
        lhs, rhs = pad_equally_0(lhs, rhs)
        t2i = self.vocabulary.token_to_id
        ids = [t2i[a + b] for a, b in zip(reversed(lhs), reversed(rhs))]
        ids.append(t2i["="])
        return ids

    def encode(self, text: str) -> list[int]:
        cleaned = self.vocabulary.clean(text)
        if "=" not in cleaned:
            return self._encode_number(cleaned)
        addition, result = cleaned.split("=")
        return self._encode_addition(addition) + self._encode_number(result)


class Decoder:
    """Inverse of :class:`Encoder` including PAD/EOS handling (reference :250-340)."""

    def __init__(self, vocabulary: Vocabulary):
        self.vocabulary = vocabulary

    def _decode_end(self, tokens_txt: str) -> str:
        if not tokens_txt:
            return ""
        if constants.EOS_TOKEN in tokens_txt:
            number, tail = tokens_txt.split(constants.EOS_TOKEN)
            return remove_leading_zeros(number[::-1]) + constants.EOS_TOKEN + tail
        return remove_leading_zeros(tokens_txt[::-1])

    def _decode_addition(self, tokens_txt: str) -> str:
        n_pad = 0
        body = tokens_txt
        if constants.PAD_TOKEN in tokens_txt:
            pieces = tokens_txt.split(constants.PAD_TOKEN)
            n_pad, body = len(pieces) - 1, pieces[-1]
        lhs = remove_leading_zeros(body[0::2][::-1])
        rhs = remove_leading_zeros(body[1::2][::-1])
        return constants.PAD_TOKEN * n_pad + lhs + "+" + rhs + "="

    def decode(self, ids_list: list[int]) -> str:
        text = "".join(self.vocabulary.id_to_token[i] for i in ids_list)
        if "=" not in text:
            return self._decode_end(text)
        addition, result = text.split("=")
        return self._decode_addition(addition) + self._decode_end(result)


class ReversedPairingTokenizer(Tokenizer):
    """Tokenizer used by the notebooks (reference :343-380); 114 ids."""

    def __init__(self):
        self.vocabulary = Vocabulary()
        self.tokenizer_encoder = Encoder(self.vocabulary)
        self.tokenizer_decoder = Decoder(self.vocabulary)
        self.token_to_id = self.vocabulary.token_to_id
        self.ntokens = self.vocabulary.ntokens

    def encode(self, text: str) -> list[int]:
        return self.tokenizer_encoder.encode(text)

    def decode(self, encoded_tokens: list[int]) -> str:
        return self.tokenizer_decoder.decode(encoded_tokens)


class CharVocabTokenizer(Tokenizer):
    """Plain character tokenizer over an arbitrary alphabet + PAD + EOS.

    Not in the reference: it is the 67-symbol tokenizer SURVEY §8d-C1 defines for
    the Shakespeare-window pool (65 distinct characters of the text, sorted, then
    PAD, EOS).  It subclasses the reference's ``Tokenizer`` ABC and exposes the
    attributes ``preprocessing.pad`` needs (``token_to_id``, ``ntokens``).
    """

    def __init__(self, alphabet: str):
        chars = sorted(set(alphabet))
        self.vocab = chars + [constants.PAD_TOKEN, constants.EOS_TOKEN]
        self.token_to_id = {tok: i for i, tok in enumerate(self.vocab)}
        self.id_to_token = dict(enumerate(self.vocab))
        self.ntokens = len(self.vocab)

    def encode(self, text: str) -> list[int]:
        t2i = self.token_to_id
        return [t2i[ch] for ch in text if ch in t2i]

    def decode(self, token_list: list[int]) -> str:
        return "".join(self.id_to_token[i] for i in token_list)
