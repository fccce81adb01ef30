"""CPU: the index arithmetic of the synthesis kernel's row classes (uno_b200/csrc/tc_rowgemm.cuh rowgemm_row / rowgemm_shift and the
host rule in backend_cuda.cu try_tc_rowgemm), restated in Python: the tiles partition the rows, a CTA only ever sees one class, and
with the class shift every lane quad (eight consecutive accumulator columns) starts on a 32-byte boundary of the output row."""
import math

import pytest


def host_rule(ldc, N, M):
    """nclass, shift_mul as try_tc_rowgemm picks them (C 32-byte aligned, switch rowgemm_parity = 1)"""
    n_tiles = (N + 255) // 256
    N_t = (((N + n_tiles - 1) // n_tiles) + 15) // 16 * 16
    parity = (ldc & 1) and n_tiles * N_t > N
    nclass, shift_mul = (2, 1) if parity else (1, 1)
    g8 = math.gcd(ldc, 8)
    if g8 < 8 and n_tiles * N_t >= N + 7 and M >= 128 * (8 // g8):
        nclass, shift_mul = 8 // g8, ldc & 7
    return nclass, shift_mul, n_tiles, N_t


def row_of(tile, i, nclass):
    cs = nclass.bit_length() - 1
    return ((tile >> cs) << (7 + cs)) + (i << cs) + (tile & (nclass - 1))


def shift_of(tile, nclass, shift_mul):
    return ((tile & (nclass - 1)) * shift_mul) & 7 if nclass > 1 else 0


@pytest.mark.parametrize("ldc", [481, 83, 301, 45, 62, 446, 54, 52, 124, 240, 223])
@pytest.mark.parametrize("M", [1408, 1024, 130, 5000])
def test_row_classes_partition_rows_and_align_quads(ldc, M):
    N = ldc
    nclass, shift_mul, n_tiles, N_t = host_rule(ldc, N, M)
    m_tiles = nclass * ((M + 128 * nclass - 1) // (128 * nclass))
    seen = set()
    for tile in range(m_tiles):
        s = shift_of(tile, nclass, shift_mul)
        assert 0 <= s <= 7 and n_tiles * N_t >= N + s            # the shifted columns fit the column tiles
        for i in range(128):
            r = row_of(tile, i, nclass)
            assert r not in seen
            seen.add(r)
            assert r % nclass == tile % nclass                    # one class per tile
            if r < M and nclass * math.gcd(ldc, 8) == 8:          # sector classes: every quad of the row on a 32-byte boundary
                assert (r * ldc - s) % 8 == 0
            if r < M and nclass == 2 and shift_mul == 1:          # parity fallback: every pair on an 8-byte boundary
                assert (r * ldc - s) % 2 == 0
    assert set(range(M)) <= seen
    # a grid that is a multiple of nclass keeps tile % nclass == block % nclass along the stride
    for gx in (nclass, 18 * nclass, (148 // n_tiles) & ~(nclass - 1)):
        if gx:
            for b in range(min(gx, 16)):
                assert all((t % nclass) == (b % nclass) for t in range(b, m_tiles, gx))
