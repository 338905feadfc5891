"""Inputs of the SURVEY §8c vectors V1-V11 (shared by CPU and GPU parity tests)."""
import numpy as np

SENTENCE = (b"He served fire and smoke; these denizens of the fields served vegetation, weather, "
            b"frost, and sun.")


def splitmix64(seed, n):
    k = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def popcount16(x):
    x = x.astype(np.uint64)
    c = np.zeros(x.shape, dtype=np.uint64)
    for b in range(16):
        c += (x >> np.uint64(b)) & np.uint64(1)
    return c


def vector_input(name):
    if name == "V1":
        return b""
    if name == "V2":
        return b"a"
    if name == "V3":
        return b"hello world"
    if name == "V4":
        return b"aaaa"
    if name == "V5":
        return bytes(1000)
    if name == "V6":
        return SENTENCE
    if name == "V7":
        return b"abcdefg" * 1000
    if name == "V8":
        return (splitmix64(1, 250000) & np.uint64(0xFF)).astype(np.uint8).tobytes()
    if name == "V9":
        return (popcount16(splitmix64(2, 250000) & np.uint64(0xFFFF)) + np.uint64(97)).astype(np.uint8).tobytes()
    if name == "V10":
        return bytes(2000000)
    if name == "V11":
        return b"aaaab" * 40000
    raise KeyError(name)
