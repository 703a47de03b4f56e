"""ORACLE (test infrastructure): restatement of the reference's PHOSC label generators
(ResPhoSCNetZSL/modules/utils/phos_generator.py:59-78 and phoc_generator.py:17-90, 'eng' version) in numpy integers.

PHOS: 11 shape counts for the whole word and for the segments of the 2-, 3-, 4- and 5-way splits (15 segments, 165 values).
PHOC: 36 presence bits [0-9a-z] of the lower-cased word for the segments of the 2..5-way splits (14 segments, 504 bits) plus
the 2 x 50 'frequent bigram' bits -- which the reference always leaves at zero, because generate_50 looks single characters
up in a list of two-letter strings (phoc_generator.py:62-69).  PHOSC = PHOS ++ PHOC (769 values; the UNet takes them as
token ids, unetPhosc.py:1124).  Pinned by tests/test_oracle_golden.py against labels produced by the reference functions."""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, "phos_alphabet.json")) as _f:
    _A = json.load(_f)
LETTERS = _A["letters"]
COUNTS = np.array(_A["counts"], dtype=np.int64)  # [52, 11]
_IDX = {c: i for i, c in enumerate(LETTERS)}
PHOS_LEN, PHOC_LEN = 165, 604


def segments(L):
    """(start, end) of every segment in reference order: whole word, then the 2..5-way splits (phos_generator.py:62-69)."""
    segs = [(0, L)]
    for split in range(2, 6):
        parts = L // split
        for mul in range(split - 1):
            segs.append((mul * parts, mul * parts + parts))
        segs.append(((split - 1) * parts, L))
    return segs


def phos(word):
    out = []
    for a, b in segments(len(word)):
        v = np.zeros(COUNTS.shape[1], dtype=np.int64)
        for ch in word[a:b]:
            v += COUNTS[_IDX[ch]]  # KeyError for a character outside a-zA-Z, as in the reference
        out.append(v)
    return np.concatenate(out)


def phoc(word):
    word = word.lower()
    out = []
    for a, b in segments(len(word))[1:]:  # no level-1 segment (phoc_generator.py:76-80)
        v = np.zeros(36, dtype=np.int64)
        for ch in word[a:b]:
            if ch.isdigit():
                v[ord(ch) - ord("0")] = 1
            elif ch.isalpha():
                v[10 + ord(ch) - ord("a")] = 1
        out.append(v)
    out.append(np.zeros(100, dtype=np.int64))  # the bigram part is never set by the reference
    return np.concatenate(out)


def phosc(word):
    w = word.replace(" ", "").replace("_", "")  # trainGWModifyCondition.py:394
    return np.concatenate([phos(w), phoc(w)])
