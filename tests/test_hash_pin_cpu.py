"""Pins the oracle's FxHasher restatement against a REAL rustc-hash 2.x build executed here.

The reference hashes k-mers with `rustc_hash::FxHasher` over `Vec<u8>` (hasher.rs:12-21, hash_iter.rs:31-45); the
crate is not vendored with the reference and no Rust toolchain exists in this image, so the reference's own hash
values cannot be produced.  But the image ships `outlines_core` (a Rust extension) whose `Vocabulary` keeps its
tokens in a `rustc_hash::FxHashMap<Vec<u8>, Vec<u32>>` and serialises it in iteration order (`__reduce__`).  A
hashbrown table iterates in bucket order and a key's bucket is `hash & (buckets - 1)` (first free slot, cyclically,
for tables narrower than one SIMD group), so the order of a tiny vocabulary is a function of the low bits of
`<Vec<u8> as Hash>::hash` under `FxHasher::default()`: write_usize(len) [length prefix], write(bytes) =
add_to_hash(hash_bytes(bytes)), finish() = rotate_left(26).  Every trial must reproduce the observed order exactly;
a wrong multiplier, seed, hash_bytes branch, missing length prefix or the rotate of rustc-hash < 2.1.1 (20) all
drop the agreement to chance (1/6 for three keys).  Key lengths cover all hash_bytes branches (n < 4, 4..7, 8..16,
17..32, > 32) and in particular the k-mer lengths the kernels specialise (17..32)."""
import random

import pytest

oc = pytest.importorskip("outlines_core")

M64 = (1 << 64) - 1
FX_K = 0xF1357AEA2E62A9C5


def _fx_vec_u8(oracle, b: bytes, rot: int) -> int:
    s = ((0 + len(b)) * FX_K) & M64                      # write_length_prefix -> write_usize(len)
    s = ((s + oracle.hash_bytes(b)) * FX_K) & M64        # write(bytes) -> add_to_hash(hash_bytes(bytes))
    return ((s << rot) | (s >> (64 - rot))) & M64        # finish()


def _place(slots, h):
    pos = h & (len(slots) - 1)
    while slots[pos] is not None:  # tables narrower than one group: first free slot, cyclically
        pos = (pos + 1) % len(slots)
    return pos


def _predicted(oracle, keys, rot):
    """hashbrown: 0 -> 4 buckets at the first insert (capacity 3), 4 -> 8 buckets at the fourth (capacity 7);
    a resize re-inserts the old table's entries in its iteration (bucket) order."""
    slots = [None] * 4
    for i, k in enumerate(keys):
        if i == 3:
            old = [x for x in slots if x is not None]
            slots = [None] * 8
            for x in old:
                slots[_place(slots, _fx_vec_u8(oracle, x, rot))] = x
        slots[_place(slots, _fx_vec_u8(oracle, k, rot))] = k
    return [x for x in slots if x is not None]


def _observed(keys):
    v = oc.Vocabulary(250, {})
    for i, k in enumerate(keys):
        v.insert(k, i)
    blob = v.__reduce__()[1][0]  # eos, n, then per entry: len, bytes, 1, id (all < 251: one byte each)
    p, out = 2, []
    assert blob[0] == 250 and blob[1] == len(keys)
    for _ in keys:
        n = blob[p]
        out.append(bytes(blob[p + 1:p + 1 + n]))
        assert blob[p + 1 + n] == 1
        p += n + 3
    return out


def _keys(rng, n, length, alphabet):
    out = []
    while len(out) < n:
        k = bytes(rng.choice(alphabet) for _ in range(length))
        if k not in out:
            out.append(k)
    return out


LENGTHS = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 18, 19, 20, 21, 24, 25, 31, 32, 33, 40, 48, 49, 64, 100]


@pytest.mark.parametrize("n_keys", [3, 7])
def test_fxhasher_restatement_matches_real_rustc_hash(oracle, n_keys):
    rng = random.Random(20260 + n_keys)
    trials = agree26 = agree20 = 0
    for length in LENGTHS:
        for alphabet in (b"ACGT", bytes(range(256))):
            for _ in range(6):
                if length == 1 and alphabet == b"ACGT" and n_keys > 4:
                    continue  # only 4 distinct keys exist
                keys = _keys(rng, n_keys, length, alphabet)
                seen = _observed(keys)
                trials += 1
                agree26 += seen == _predicted(oracle, keys, 26)
                agree20 += seen == _predicted(oracle, keys, 20)
    assert trials >= 280
    assert agree26 == trials, f"rotate 26: {agree26}/{trials}"
    assert agree20 < trials // 2, f"rotate 20 should be at chance level, got {agree20}/{trials}"


def _capacity(buckets):
    return buckets - 1 if buckets < 8 else buckets // 8 * 7


def _buckets_for(cap):
    if cap < 8:
        return 4 if cap < 4 else 8
    want, p = -(-cap * 8 // 7), 1
    while p < want:
        p *= 2
    return p


def _place_wide(slots, h):
    """hashbrown's probe: the first free slot in the 16 control bytes starting at hash & mask, else the next window"""
    n, pos, stride = len(slots), h & (len(slots) - 1), 0
    while True:
        for i in range(min(16, n)):
            j = (pos + i) & (n - 1)
            if slots[j] is None:
                return j
        stride += 16
        pos = (pos + stride) & (n - 1)


@pytest.mark.parametrize("n_keys,length", [(14, 20), (56, 21), (100, 31), (200, 20), (200, 33)])
def test_fxhasher_low_byte_in_wide_tables(oracle, n_keys, length):
    """Up to 200 keys -> 256 buckets: the iteration order now depends on the low 8 bits of every key's hash
    (growth 4 -> 8 -> 16 -> ... re-inserts in bucket order); it must be predicted exactly."""
    rng = random.Random(n_keys * 100 + length)
    for _ in range(5):
        keys = _keys(rng, n_keys, length, bytes(range(256)))
        slots, items = [], 0
        for k in keys:
            if not slots or items == _capacity(len(slots)):
                old = [x for x in slots if x is not None]
                slots = [None] * _buckets_for(max(items + 1, _capacity(len(slots)) + 1 if slots else 1))
                for x in old:
                    slots[_place_wide(slots, _fx_vec_u8(oracle, x, 26))] = x
            slots[_place_wide(slots, _fx_vec_u8(oracle, k, 26))] = k
            items += 1
        v = oc.Vocabulary(250, {})
        for i, k in enumerate(keys):
            v.insert(k, i % 200)
        blob = v.__reduce__()[1][0]
        p = 2
        if blob[1] == 251:  # bincode varint: u16 follows
            assert blob[2] | (blob[3] << 8) == n_keys
            p = 4
        seen = []
        for _ in keys:
            seen.append(bytes(blob[p + 1:p + 1 + length]))
            p += length + 3
        assert seen == [x for x in slots if x is not None]


def test_pure_python_restatement_agrees(oracle):
    """the independent pure-Python hash_bytes (tests/test_oracle_cpu.py) and the C oracle agree on the same keys"""
    from tests.test_oracle_cpu import py_hash_bytes
    rng = random.Random(5)
    for length in LENGTHS:
        k = bytes(rng.randrange(256) for _ in range(length))
        assert py_hash_bytes(k) == oracle.hash_bytes(k)
