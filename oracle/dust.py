"""Oracle (test infrastructure): symmetric DUST low-complexity soft-masking.

The reference calls `pydustmasker.DustMasker(seq, window_size=64, score_threshold=20).mask()`
(seqops/io.py:105-108).  pydustmasker (>= 1.0.3, Rust) is NOT vendored by the reference and not
installable here; it implements the SDUST algorithm (Morgulis et al. 2006, "A fast and symmetric
DUST implementation to mask low-complexity DNA sequences"), of which the canonical
implementation is `sdust.c` in minimap2.  This file restates that published algorithm:

  * words are triplets (64 possible); a window holds at most W - 2 words (W = 64 bases);
  * the score of an interval with word counts c_t is  sum_t c_t (c_t - 1) / 2  over (len - 1);
    intervals scoring above T/10 (T = 20) with no higher-scoring sub-interval are "perfect";
  * every perfect interval of every window is masked; bases other than A/C/G/T break the
    sequence into independent pieces.

PARITY UNPINNED: no fixture of pydustmasker output exists (the reference tests only import it).
`mask()` returns the sequence with masked bases lower-cased, like the reference consumes it.
"""
from __future__ import annotations

W_LEN = 3
W_TOT = 64


def sdust_intervals(seq: str, W: int = 64, T: int = 20) -> list[tuple[int, int]]:
    """Masked half-open intervals [start, finish) in sequence coordinates."""
    code = {"A": 0, "C": 1, "G": 2, "T": 3, "a": 0, "c": 1, "g": 2, "t": 3}
    res: list[list[int]] = []
    P: list[list[int]] = []          # [start, finish, r, l], descending start then ascending finish
    w: list[int] = []                # the window's words (deque)
    cv = [0] * W_TOT
    cw = [0] * W_TOT
    rv = rw = L = 0

    def save_masked(start):
        nonlocal P
        if not P or P[-1][0] >= start:
            return
        p = P[-1]
        saved = False
        if res:
            s, f = res[-1]
            if p[0] <= f:
                res[-1][1] = max(f, p[1])
                saved = True
        if not saved:
            res.append([p[0], p[1]])
        i = len(P) - 1
        while i >= 0 and P[i][0] < start:
            i -= 1
        P = P[:i + 1]

    def shift_window(t):
        nonlocal rv, rw, L
        if len(w) >= W - W_LEN + 1:
            s = w.pop(0)
            cw[s] -= 1
            rw -= cw[s]
            if L > len(w):
                L -= 1
                cv[s] -= 1
                rv -= cv[s]
        w.append(t)
        L += 1
        rw += cw[t]
        cw[t] += 1
        rv += cv[t]
        cv[t] += 1
        if cv[t] * 10 > T * 2:
            while True:
                s = w[len(w) - L]
                cv[s] -= 1
                rv -= cv[s]
                L -= 1
                if s == t:
                    break

    def find_perfect(start):
        c = cv[:]
        r = rv
        max_r = max_l = 0
        for i in range(len(w) - L - 1, -1, -1):
            t = w[i]
            r += c[t]
            c[t] += 1
            new_r, new_l = r, len(w) - i - 1
            if new_r * 10 > T * new_l:
                j = 0
                while j < len(P) and P[j][0] >= i + start:
                    p = P[j]
                    if max_r == 0 or p[2] * max_l > max_r * p[3]:
                        max_r, max_l = p[2], p[3]
                    j += 1
                if max_r == 0 or new_r * max_l >= max_r * new_l:
                    max_r, max_l = new_r, new_l
                    P.insert(j, [i + start, len(w) + (W_LEN - 1) + start, new_r, new_l])

    n = len(seq)
    l = t = 0
    for i in range(n + 1):
        b = code.get(seq[i], 4) if i < n else 4
        if b < 4:
            l += 1
            t = ((t << 2) | b) & (W_TOT - 1)
            if l >= W_LEN:
                start = max(l - W, 0) + (i + 1 - l)
                save_masked(start)
                shift_window(t)
                if rw * 10 > L * T:
                    find_perfect(start)
        else:
            start = max(l - W + 1, 0) + (i + 1 - l)
            while P:
                save_masked(start)
                start += 1
            l = t = 0
            w.clear()
            for k in range(W_TOT):
                cv[k] = cw[k] = 0
            rv = rw = L = 0
    return [(s, f) for s, f in res]


def mask(seq: str, W: int = 64, T: int = 20) -> str:
    out = list(seq)
    for s, f in sdust_intervals(seq, W, T):
        for i in range(s, min(f, len(out))):
            out[i] = out[i].lower()
    return "".join(out)


def mask_bits(seq: str, W: int = 64, T: int = 20):
    import numpy as np
    m = np.zeros(len(seq), dtype=bool)
    for s, f in sdust_intervals(seq, W, T):
        m[s:f] = True
    return m
