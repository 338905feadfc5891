"""CPU check of the two arguments the BWT sort kernels rest on beyond plain prefix doubling
(csrc/bwt_common.cuh), against a naive cyclic rotation sort with the reference's tie rule
(equal rotations by descending start index, lib/bwt.rs:733-749):

* Period: inside a run S[x] == S[x+p] (s <= x < e-p) any two rotations a < b of [s, e) with
  a == b (mod p) are ordered the same way, namely like the rotations e-p and e.
* KeyCode: ranking rotations by the packed round-0 key (k symbols base sigma + a coarse next symbol)
  is consistent with the true order, and equal keys imply equal k-symbol prefixes.
"""
import random

import numpy as np


def _sorted_rotations(S):
    n = len(S)
    return sorted(range(n), key=lambda i: (S[i:] + S[:i], -i))


def _period_run(S, W=4):
    """what detect_period computes (window W instead of 32 bytes): (p, s, e, asc) or None"""
    n = len(S)
    m = n // 2
    for p in range(1, n // 4):
        if S[m:m + W] != S[m + p:m + p + W]:
            continue
        viol = [x for x in range(n - p) if S[x] != S[x + p]]
        s = max([x + 1 for x in viol if x < m], default=0)
        e = min([x for x in viol if x >= m], default=n - p) + p
        a, b = (e - p) % n, e % n
        ra, rb = S[a:] + S[:a], S[b:] + S[:b]
        if ra == rb:
            return None                      # cyclically p-periodic: the tie rule decides
        return p, s, e, ra < rb
    return None


def test_same_class_rotations_of_a_run_are_ordered_by_index():
    random.seed(1)
    checked = 0
    for _ in range(1500):
        p = random.randint(1, 5)
        unit = "".join(random.choice("ab") for _ in range(p))
        n = random.randint(16, 60)
        S = list((unit * n)[:n])
        for _ in range(random.randint(0, 2)):                       # damage
            S[random.randrange(n)] = random.choice("abc")
        S = "".join(random.choice("abc") for _ in range(random.randint(0, 3))) + "".join(S)
        run = _period_run(S)
        if run is None:
            continue
        p, s, e, asc = run
        pos = {i: k for k, i in enumerate(_sorted_rotations(S))}
        for i in range(s, e):
            for j in range(i + p, e, p):
                assert (pos[i] < pos[j]) == asc, (S, p, s, e, i, j, asc)
                checked += 1
    assert checked > 10000


def test_packed_round0_key_is_consistent_with_the_rotation_order():
    rng = np.random.default_rng(3)
    for sigma_target in (2, 3, 17, 33, 53, 56, 90, 101, 102, 200):
        n = 400
        alphabet = np.sort(rng.choice(256, size=sigma_target, replace=False))
        S = alphabet[rng.integers(0, sigma_target, n)]
        S[: sigma_target] = alphabet                                 # every symbol present
        syms = np.unique(S)
        sigma = len(syms)
        k = 5 if sigma > 101 else 6 if sigma > 52 else 7 if sigma > 32 else 8
        L = min(sigma, (1 << 40) // sigma ** k)
        code = {int(b): i for i, b in enumerate(syms)}
        keys = []
        for i in range(n):
            if k == 5 and sigma > 101:
                key = 0
                for j in range(5):
                    key = (key << 8) | int(S[(i + j) % n])
            else:
                key = 0
                for j in range(k):
                    key = key * sigma + code[int(S[(i + j) % n])]
                key = key * L + (code[int(S[(i + k) % n])] * L) // sigma
            assert key < (1 << 40)
            keys.append(key)
        raw = bytes(S.tolist())
        order = _sorted_rotations(raw)
        for a, b in zip(order, order[1:]):
            assert keys[a] <= keys[b]                                # consistent with the true order
        for a in range(n):
            for b in range(a + 1, n):
                if keys[a] == keys[b]:                               # ties imply equal k-prefixes
                    assert all(S[(a + j) % n] == S[(b + j) % n] for j in range(k))
