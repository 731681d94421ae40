"""Shared case generators for the parity tests.

``REF_KATS`` restates every known-answer vector of the reference's own unit tests
(/root/reference/tests/test_monkey_moore.cpp, cited per entry) so the oracle and the
CUDA path can both be pinned against them without the reference being present.
"""
import numpy as np

HIRAGANA = "あいうえおかきくけこさしすせそたちつてとなにぬねのはひふへほまみむめもやゆよらりるれろわをゃっゅょ"
# src/gui/constants.hpp:48 (MM_DEFAULT_HIRAGANA) -- the same 49 characters
KATAKANA = "アイウエオカキクケコサシスセソタチツテトナニヌネノハヒフヘホマミムメモヤユヨラリルレロワヲャッュョ"
VOWEL_SEQ = "aiueobcdfghjklmnpqrstvwxyz"


def shift_alpha(seq, lower_shift, upper_shift, dtype):
    """tests/common.hpp:110-125 shift_alpha_values"""
    out = []
    for c in seq:
        v = ord(c) if isinstance(c, str) else int(c)
        if ord("a") <= v <= ord("z"):
            v += lower_shift
        elif ord("A") <= v <= ord("Z"):
            v += upper_shift
        out.append(v & (0xFF if dtype == np.uint8 else 0xFFFF))
    return np.array(out, dtype=dtype)


def _chars(s):
    return [c if isinstance(c, int) else ord(c) for c in s]


def ascii_map(a, A, bits):
    m = (1 << bits) - 1
    return {ord("A"): A & m, ord("a"): a & m}


def seq_map(seq, first_value):
    return {ord(c): first_value + i for i, c in enumerate(seq)}


def ref_kats():
    """-> list of dicts: name, bits, data, pattern kwargs, expected positions, expected maps (or None)"""
    K = []
    # test_monkey_moore.cpp:16-35
    d = shift_alpha("dddccacatchaat", 3, 3, np.uint8)
    K.append(dict(name="nowc8-catch", bits=8, data=d, keyword="catch", pos=[6],
                  maps=[ascii_map(ord("a") + 3, ord("A") + 3, 8)]))
    K.append(dict(name="nowc8-maca", bits=8, data=d, keyword="maca", pos=[], maps=[]))
    # :38-52
    d = np.array(_chars("auqqtkcaoaugka"), dtype=np.uint8)
    K.append(dict(name="nowc8-seq-match", bits=8, data=d, keyword="match", char_seq=VOWEL_SEQ, pos=[8],
                  maps=[{ord(c): ord("a") + i for i, c in enumerate(VOWEL_SEQ)}]))
    # :56-78
    d = shift_alpha(["q", "u", "e", "s", "t", "i", "o", "n", " ", "o", "f", " ", "p", "r", "i", "c",
                     "e", 0, "t", "h", "e", " ", "l", "a", "s", "t", " ", "w", "i", "s", "h", 0], -16, -16, np.uint16)
    K.append(dict(name="nowc16-price", bits=16, data=d, keyword="price", pos=[12],
                  maps=[ascii_map(ord("a") - 16, ord("A") - 16, 16)]))
    K.append(dict(name="nowc16-station", bits=16, data=d, keyword="station", pos=[], maps=[]))
    # :81-104
    d = np.array([1, 12, 16, 110, 44, 16, 12, 16, 17, 26, 110, 22, 44, 22, 110, 26,
                  21, 45, 110, 31, 7, 31, 13], dtype=np.uint16)
    K.append(dict(name="nowc16-hiragana", bits=16, data=d, keyword="わたしたちは", char_seq=HIRAGANA, pos=[4],
                  maps=[seq_map(HIRAGANA, 1)]))
    # :110-146
    d = shift_alpha("thebittertasteoflemonwithbutter,", 8, 8, np.uint8)
    m8 = ascii_map(ord("a") + 8, ord("A") + 8, 8)
    K.append(dict(name="wc8-b*tter", bits=8, data=d, keyword="b*tter", wildcard=ord("*"), pos=[3, 25], maps=[m8, m8]))
    K.append(dict(name="wc8-t?ste", bits=8, data=d, keyword="t?ste", wildcard=ord("?"), pos=[9], maps=[m8]))
    K.append(dict(name="wc8-past*-literal", bits=8, data=d, keyword="past*", pos=[], maps=[]))
    # :148-174
    d = shift_alpha("TheBitterTruthAboutBetterButter.", -32, 24, np.uint8)
    mm = ascii_map(ord("a") - 32, ord("A") + 24, 8)
    K.append(dict(name="wc8-mixed-B*tter", bits=8, data=d, keyword="B*tter", wildcard=ord("*"), pos=[3, 19, 25],
                  maps=[mm, mm, mm]))
    K.append(dict(name="wc8-mixed-Matter", bits=8, data=d, keyword="Matter", pos=[], maps=[]))
    # :177-191
    d = np.array(_chars("auqqtkcaoaugka"), dtype=np.uint8)
    K.append(dict(name="wc8-seq-*at*h", bits=8, data=d, keyword="*at*h", wildcard=ord("*"), char_seq=VOWEL_SEQ,
                  pos=[8], maps=[{ord(c): ord("a") + i for i, c in enumerate(VOWEL_SEQ)}]))
    # :195-220
    d = shift_alpha("They muttered: Butter, BETTER, Butcher, matter", 15, -9, np.uint16)
    K.append(dict(name="wc16-But**er", bits=16, data=d, keyword="But**er", wildcard=ord("*"), pos=[31],
                  maps=[ascii_map(ord("a") + 15, ord("A") - 9, 16)]))
    K.append(dict(name="wc16-*ITTER", bits=16, data=d, keyword="*ITTER", wildcard=ord("*"), pos=[], maps=[]))
    # :223-246
    seq = HIRAGANA + "学校行"
    d = np.array([1, 12, 16, 26, 111, 50, 51, 22, 111, 52, 7, 31, 13, 6, 112, 111,
                  44, 16, 12, 35, 111, 52, 7, 16, 2, 113], dtype=np.uint16)
    K.append(dict(name="wc16-kanji", bits=16, data=d, keyword="**に*行きますか", wildcard=ord("*"), char_seq=seq,
                  pos=[5], maps=[seq_map(seq, 1)]))
    # :251-273
    d = np.array([0x00, 0x00, 0x25, 0x26, 0x25, 0x26, 0x27, 0x28, 0x29, 0x30, 0x20, 0x20, 0x00, 0x00, 0x01, 0x00,
                  0x01, 0x00, 0x00, 0x89, 0x00, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x81, 0x00, 0x00, 0x01, 0x00, 0x00],
                 dtype=np.uint8)
    K.append(dict(name="vs8-hit", bits=8, data=d, values=[60, 61, 62, 63, 64, 71], pos=[4, 21], maps=[{}, {}]))
    K.append(dict(name="vs8-miss", bits=8, data=d, values=[80, 81, 82, 83, 84, 85, 86], pos=[], maps=[]))
    # :276-300
    d = np.array([0x0000, 0x0100, 0x0135, 0x0136, 0x0135, 0x0136, 0x0137, 0x0138,
                  0x0139, 0x0140, 0x0120, 0x0120, 0x0000, 0x0100, 0x0101, 0x0000,
                  0x0101, 0x0089, 0x0000, 0x0045, 0x0046, 0x0047, 0x0048, 0x0049,
                  0x0050, 0x0000, 0x0100, 0x0000, 0x0100, 0x0001, 0x0100, 0x0000], dtype=np.uint16)
    K.append(dict(name="vs16-hit", bits=16, data=d, values=[105, 106, 107, 108, 109, 116], pos=[4, 19], maps=[{}, {}]))
    K.append(dict(name="vs16-miss", bits=16, data=d, values=[200, 201, 205, 208, 209], pos=[], maps=[]))
    # :304-344 (skip-table regression)
    d = np.array([0x98, 0x94, 0x00, 0xFF, 0xFF, 0x00, 0x01, 0xA5, 0xA1, 0x94, 0x85, 0x98, 0x94], dtype=np.uint8)
    K.append(dict(name="reg8-0xFF", bits=8, data=d, keyword="text", pos=[9], maps=None))
    d = np.array([0x1098, 0x1094, 0x0000, 0xFFFF, 0xFFFF, 0x1000, 0x1001, 0x10A5,
                  0x10A1, 0x1094, 0x1085, 0x1098, 0x1094], dtype=np.uint16)
    K.append(dict(name="reg16-0xFFFF", bits=16, data=d, keyword="text", pos=[9], maps=None))
    return K


# /root/reference/tests/test_search_engine.cpp:26-81
ENGINE8_FILE = np.array([
    0x94, 0x85, 0x98, 0x94, 0x10, 0x10, 0x11, 0x11,
    0x00, 0x94, 0x85, 0x98, 0x94, 0x00, 0xFF, 0xFF,
    0x00, 0x00, 0x01, 0x0A, 0xFF, 0xFF, 0x00, 0x00,
    0x00, 0x94, 0x85, 0x94, 0x85, 0x98, 0x94, 0x00,
    0xFF, 0x00, 0x0A, 0xFF, 0xFF, 0x01, 0x00, 0x00,
    0xFF, 0x00, 0x0A, 0xFF, 0xFF, 0x01, 0x00, 0x00,
    0x00, 0xFF, 0x94, 0x85, 0x98, 0x94, 0x00, 0xFF,
    0x00, 0x01, 0xA5, 0xA1, 0x94, 0x85, 0x98, 0x94], dtype=np.uint8)
ENGINE8_OFFSETS = [0, 9, 27, 50, 60]
ENGINE8_BLOCKS = [128, 8, 23, 29]

# :83-137
ENGINE16_FILE = np.array([
    0x1094, 0x1085, 0x1098, 0x1094, 0x0010, 0x0010, 0x0011, 0x0011,
    0x0000, 0x1094, 0x1085, 0x1098, 0x1094, 0x0000, 0xFFFF, 0xFFFF,
    0x0000, 0x0000, 0x0001, 0x000A, 0xFFFF, 0xFFFF, 0x0000, 0x0000,
    0x0000, 0x1094, 0x1085, 0x1094, 0x1085, 0x1098, 0x1094, 0x0000,
    0xFFFF, 0x0000, 0x000A, 0xFFFF, 0xFFFF, 0x0001, 0x0000, 0x0000,
    0xFFFF, 0x0000, 0x000A, 0xFFFF, 0xFFFF, 0x0001, 0x0000, 0x0000,
    0x0000, 0xFFFF, 0x1094, 0x1085, 0x1098, 0x1094, 0x0000, 0x00FF,
    0x0000, 0x0110, 0xA510, 0x01A1, 0x1094, 0x1085, 0x1098, 0x1094], dtype=np.uint16)
ENGINE16_OFFSETS = [0, 18, 54, 100, 120]
ENGINE16_BLOCKS_LE = [256, 16, 47, 58]
ENGINE16_BLOCKS_BE = [512, 24, 47, 58]


# ----------------------------------------------------------------------------
# randomised patterns / data for differential testing
# ----------------------------------------------------------------------------

def random_pattern(rng, bits):
    """-> kwargs for Oracle/Ref/product (keyword|values, wildcard, char_seq)"""
    kind = rng.choice(["ascii", "ascii", "wild", "mixed", "seq", "seqwild", "values", "raw"])
    L = int(rng.integers(2, 13))
    low = "abcdefghijklmnopqrstuvwxyz"
    up = low.upper()
    if kind == "values":
        span = int(rng.choice([3, 20, 200, 400]))
        base = int(rng.integers(-50, 300))
        return dict(values=[int(base + rng.integers(-span, span + 1)) for _ in range(L)])
    if kind == "raw":
        # arbitrary small code points (incl. non letters, repeats)
        alpha = int(rng.choice([2, 3, 8, 64]))
        base = int(rng.integers(1, 200))
        return dict(keyword=[base + int(rng.integers(0, alpha)) for _ in range(L)], wildcard=0)
    if kind == "ascii":
        alpha = low[: int(rng.choice([2, 3, 5, 26]))]
        return dict(keyword="".join(rng.choice(list(alpha)) for _ in range(L)), wildcard=0)
    if kind == "wild":
        alpha = low[: int(rng.choice([2, 3, 5, 26]))]
        kw = [rng.choice(list(alpha)) for _ in range(L)]
        for _ in range(int(rng.integers(1, max(2, L // 2)))):
            kw[int(rng.integers(0, L))] = "*"
        return dict(keyword="".join(kw), wildcard=ord("*"))
    if kind == "mixed":
        n = int(rng.choice([2, 3, 26]))
        kw = [rng.choice(list(low[:n] + up[:n])) for _ in range(L)]
        if rng.random() < 0.5:
            kw[int(rng.integers(0, L))] = "*"
        if rng.random() < 0.3:
            kw[int(rng.integers(0, L))] = rng.choice(list(" .1"))
        return dict(keyword="".join(kw), wildcard=ord("*"))
    seq = HIRAGANA if rng.random() < 0.5 else VOWEL_SEQ
    n = int(rng.choice([2, 4, len(seq)]))
    kw = [rng.choice(list(seq[:n])) for _ in range(L)]
    if rng.random() < 0.15:
        kw[int(rng.integers(0, L))] = "Z" if seq is HIRAGANA else "あ"  # not in the sequence -> index 0
    if kind == "seqwild":
        for _ in range(int(rng.integers(1, max(2, L // 2)))):
            kw[int(rng.integers(0, L))] = "*"
        return dict(keyword="".join(kw), wildcard=ord("*"), char_seq=seq)
    return dict(keyword="".join(kw), wildcard=0, char_seq=seq)


def pattern_values(pat):
    """Integer "letter values" of a pattern, for planting matches: list of (int | None for wildcard)."""
    if "values" in pat and pat["values"] is not None:
        return list(pat["values"])
    kw = pat["keyword"]
    cps = [c if isinstance(c, int) else ord(c) for c in kw]
    wc = pat.get("wildcard", 0)
    seq = pat.get("char_seq") or ""
    out = []
    for c in cps:
        if c == wc:
            out.append(None)
        elif seq:
            idx = {ord(ch): i for i, ch in enumerate(seq)}
            out.append(idx.get(c, 0))
        else:
            out.append(c)
    return out


def random_data(rng, bits, n, pat=None):
    """Low/high entropy element streams, optionally with planted shifted copies of the pattern."""
    dtype = np.uint8 if bits == 8 else np.uint16
    vmax = (1 << bits) - 1
    style = rng.choice(["uniform", "small", "tiny", "walk", "const", "edge"])
    if style == "uniform":
        d = rng.integers(0, vmax + 1, n)
    elif style == "small":
        d = rng.integers(0, 16, n) + int(rng.integers(0, vmax - 16))
    elif style == "tiny":
        d = rng.integers(0, int(rng.choice([2, 3, 4])), n) + int(rng.integers(0, vmax - 4))
    elif style == "walk":
        d = np.cumsum(rng.integers(-3, 4, n)) + int(rng.integers(0, vmax))
    elif style == "const":
        d = np.full(n, int(rng.integers(0, vmax + 1)))
    else:  # values hugging 0 / max to exercise wrap-around
        d = rng.choice([0, 1, 2, vmax - 2, vmax - 1, vmax], n)
    d = (d & vmax).astype(dtype)
    if pat is not None and n > 0:
        vals = pattern_values(pat)
        L = len(vals)
        for _ in range(int(rng.integers(0, 6))):
            if n < L:
                break
            at = int(rng.integers(0, n - L + 1))
            base = int(rng.integers(0, vmax + 1))
            first = next((v for v in vals if v is not None), 0)
            for i, v in enumerate(vals):
                if v is not None:
                    d[at + i] = (base + v - first) & vmax
    return d
