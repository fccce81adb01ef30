"""Coverage check (numpy, CPU) of tc_kpipe.cuh's three loader paths for 8 and 16 loader warps: every (row, k) element of a
128 x 32 chunk is written exactly once, to the K-major interleave offset the UMMA descriptor reads (k/4 * LBO + row * 16 +
k%4 * 4), from the global element the path is meant to fetch."""
import numpy as np

kLboA = 128 * 16 + 16
kKC = 32


def want_offset(row, k):
    return (k // 4) * kLboA + row * 16 + (k % 4) * 4


def check(LW, path, lda=100, tile=3, kc=2, cls=1):
    RB, RP4 = LW * 4, 128 // LW
    RP = 128 // RB
    seen = {}
    for ltid in range(LW * 32):
        warp, lane = ltid >> 5, ltid & 31
        if path in ("vec16", "rclass"):
            kq, rbase = ltid & 7, ltid >> 3
            so = kq * kLboA + rbase * 16
            for i in range(RP):
                if path == "vec16":
                    row0 = tile * 128 + rbase
                    grow = row0 + RB * i
                    src0 = grow * lda + kc * kKC + kq * 4
                    tile_row = rbase + RB * i
                else:
                    sh = (cls * lda) & 3
                    row0 = (tile >> 2) * 512 + 4 * rbase + cls
                    grow = row0 + 4 * RB * i
                    src0 = grow * lda - sh + kc * kKC + kq * 4
                    tile_row = rbase + RB * i
                    assert src0 % 4 == 0
                for e in range(4):
                    off = so + i * (RB * 16) + 4 * e
                    assert off not in seen
                    seen[off] = (tile_row, kq * 4 + e, src0 + e, grow)
        else:
            so = (lane >> 2) * kLboA + (lane & 3) * 4 + warp * 16
            for i in range(RP4):
                grow = tile * 128 + warp + LW * i
                off = so + i * (LW * 16)
                assert off not in seen
                seen[off] = (warp + LW * i, lane, grow * lda + kc * kKC + lane, grow)
    assert len(seen) == 128 * 32
    for off, (row, k, src, grow) in seen.items():
        assert off == want_offset(row, k), (LW, path, off, row, k)
        if path == "rclass":
            sh = (cls * lda) & 3
            assert grow == (tile >> 2) * 512 + 4 * row + cls and src == grow * lda - sh + kc * kKC + k
        else:
            assert grow == tile * 128 + row and src == grow * lda + kc * kKC + k


for LW in (8, 16):
    for path in ("vec16", "scalar4", "rclass"):
        for lda in ((100, 104) if path == "vec16" else (101, 83, 130)):
            for cls in range(4):
                check(LW, path, lda=lda, tile=4 + cls, cls=cls)
print("loader paths: every chunk element written once, at the descriptor's offset, from the intended address")
