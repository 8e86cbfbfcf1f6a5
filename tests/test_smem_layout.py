"""Shared-memory fragment layout of the DGEMM kernel (jues.jl_b200/csrc/dgemm_sm100.cuh):
the (row, k) -> byte-offset maps must (a) agree with what TMA's 128-byte swizzle writes,
(b) be bijections onto the tile, (c) make every LDS.64 of a DMMA fragment bank-conflict-free
(16 lanes of a half-warp hit 16 distinct 8-byte bank pairs), for both operand layouts."""
import itertools

import pytest


def frag_off(kc, r, k):
    if kc:
        return r * 128 + ((((k >> 1) ^ (r & 7)) << 4) | ((k & 1) << 3))
    return (r >> 4) * 2048 + k * 128 + (((((r & 15) >> 1) ^ (k & 7)) << 4) | ((r & 1) << 3))


def frag_row(kc, g):
    if kc:
        return (g & 1) | ((g & 2) << 1) | ((g & 4) >> 1)
    return g


def frag_k(t, s):
    return 2 * t + ((t & 1) ^ s)


def tma_swizzle128(linear_byte):
    """Where TMA (SWIZZLE_128B) stores the byte that an unswizzled box would hold at
    `linear_byte` (box base 1024-byte aligned): bits [4,7) ^= bits [7,10)."""
    return linear_byte ^ (((linear_byte >> 7) & 7) << 4)


@pytest.mark.parametrize("R", [64, 128])
def test_offsets_match_tma_swizzle(R):
    # KC: one box {16 k, R rows}: element (r,k) at linear r*128 + k*8
    for r, k in itertools.product(range(R), range(16)):
        assert frag_off(True, r, k) == tma_swizzle128(r * 128 + k * 8)
    # RC: R/16 boxes {16 r, 16 k} of 2 KB: element (r,k) at box*2048 + k*128 + (r%16)*8
    for r, k in itertools.product(range(R), range(16)):
        lin = k * 128 + (r % 16) * 8
        assert frag_off(False, r, k) == (r // 16) * 2048 + tma_swizzle128(lin)


@pytest.mark.parametrize("kc", [True, False])
def test_bijection(kc):
    offs = {frag_off(kc, r, k) for r in range(128) for k in range(16)}
    assert len(offs) == 128 * 16
    assert min(offs) == 0 and max(offs) == 128 * 16 * 8 - 8


def test_k_and_row_maps_are_permutations():
    assert sorted(frag_k(t, s) for t in range(4) for s in range(2)) == list(range(8))
    for kc in (True, False):
        assert sorted(frag_row(kc, g) for g in range(8)) == list(range(8))


@pytest.mark.parametrize("kc", [True, False])
def test_bank_conflict_free(kc):
    for warp_row0 in (0, 16, 32, 64, 96):
        for i in range(8):                      # 8-row tiles inside a 64-row warp tile
            for kb in (0, 1):
                for s in (0, 1):
                    for half in (0, 1):
                        banks = set()
                        for lane in range(16 * half, 16 * half + 16):
                            g, t = lane >> 2, lane & 3
                            r = warp_row0 + 8 * i + frag_row(kc, g)
                            k = kb * 8 + frag_k(t, s)
                            off = frag_off(kc, r, k)
                            banks.add((off % 128) // 8)
                        assert len(banks) == 16, (kc, warp_row0, i, kb, s, half)


@pytest.mark.parametrize("kc", [True, False])
def test_tile_increment_rule(kc):
    """The kernel derives tile i's offset from tile (i&1) by adding (i>>1)*2048."""
    for g in range(8):
        for t in range(4):
            for s in (0, 1):
                for kb in (0, 1):
                    for i in range(8):
                        k = kb * 8 + frag_k(t, s)
                        full = frag_off(kc, 8 * i + frag_row(kc, g), k)
                        short = frag_off(kc, 8 * (i & 1) + frag_row(kc, g), k) + (i >> 1) * 2048
                        assert full == short
