"""Shared constants (mirror of the reference's `src/configs/constants.py:10,19-20`).

Only three values exist in the reference and the hot path needs the two
special-token spellings to stay identical so tokenizer ids line up.
"""

RANDOM_SEED = 42

PAD_TOKEN = "<PAD>"
EOS_TOKEN = "<EOS>"
