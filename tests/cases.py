"""Shape tables shared by the golden generator and the tests (kept in sync with tests/golden/make_golden.py)."""

SPECTRAL_CASES = {
    "s1d": (2, 3, 4, (32,), (24,), (5,)),
    "s2d_same": (2, 3, 5, (24, 24), (24, 24), (7, 8)),
    "s2d_down_odd": (2, 3, 4, (37, 41), (19, 23), (5, 6)),
    "s2d_up_odd": (2, 2, 3, (8, 8), (32, 31), (3, 3)),
    "s2d_overlap_out": (1, 4, 2, (16, 16), (8, 8), (6, 4)),
    "s2d_overlap_in": (1, 2, 3, (10, 10), (12, 12), (6, 5)),
    "s3d_down": (1, 2, 3, (16, 16, 13), (12, 12, 13), (4, 4, 5)),
    "s3d_up": (1, 2, 2, (8, 8, 20), (16, 16, 31), (3, 3, 7)),
    "s3d_overlap": (1, 2, 2, (16, 16, 20), (8, 8, 20), (6, 6, 7)),
}

POINTWISE_CASES = {
    "p2d_down": (2, 3, 4, (16, 16), (8, 8)),
    "p2d_up": (2, 5, 2, (8, 8), (16, 16)),
    "p2d_odd": (1, 2, 3, (37, 41), (19, 23)),
    "p2d_same": (1, 3, 3, (12, 12), (12, 12)),
    "p2d_mixed": (1, 4, 2, (12, 14), (30, 9)),
    "p3d_down": (1, 2, 3, (16, 16, 13), (12, 12, 13)),
    "p3d_up": (1, 3, 2, (8, 8, 20), (16, 16, 31)),
    "p3d_same": (1, 2, 2, (12, 12, 10), (12, 12, 10)),
}

BLOCK_CASES = {
    "b2d_down_norm": (2, 3, 4, (20, 24), (10, 12), (4, 4), True, True),
    "b2d_down": (2, 3, 4, (20, 24), (10, 12), (4, 4), False, True),
    "b2d_up": (2, 4, 2, (10, 12), (20, 24), (4, 4), False, True),
    "b2d_up_norm_lin": (2, 4, 2, (10, 12), (20, 24), (4, 4), True, False),
    "b2d_same_lin": (2, 4, 3, (10, 12), (10, 12), (4, 4), False, False),
    "b3d_down_norm": (1, 2, 3, (12, 12, 10), (8, 8, 10), (3, 3, 4), True, True),
    "b3d_up": (1, 3, 2, (8, 8, 10), (12, 12, 10), (3, 3, 4), False, True),
}

RESAMPLE_PAIRS = ((481, 240), (240, 120), (120, 240), (240, 481), (446, 223), (223, 111), (111, 223), (223, 446),
                  (64, 48), (48, 32), (32, 16), (16, 32), (32, 48), (48, 64))

# model glue (lift / project): name -> (batch, dims, pad_lo, pad_hi, raw_ch, grid_ch, hidden, out_ch)
LIFT_CASES = {
    "darcy": (2, (21, 19), (0, 0), (5, 5), 1, 2, 16, 32),          # UNO_9: right/bottom pad
    "ns2d": (2, (16, 16), (3, 3), (3, 3), 10, 4, 16, 32),           # UNO: pad on all sides
    "ns2d_nopad": (1, (17, 13), (0, 0), (0, 0), 10, 4, 16, 32),
    "ns3d": (1, (8, 8, 10), (0, 0, 0), (0, 0, 3), 1, 5, 12, 8),     # Uno3D_T10: time axis trailing pad
    "ns3d_both": (1, (6, 5, 7), (0, 0, 2), (0, 0, 2), 1, 5, 12, 8),
    "wide": (1, (9, 33), (1, 0), (0, 2), 3, 2, 32, 64),             # width 64: hidden 32
    "ragged": (3, (7, 37), (0, 0), (2, 1), 1, 2, 16, 24),           # > 1 tile with a ragged tail, out_ch not a multiple of 8
}
# name -> (batch, cropped dims, crop_lo, crop_hi, src_ch, hidden, out_ch)
PROJECT_CASES = {
    "darcy": (2, (21, 19), (0, 0), (5, 5), (32, 32), 32, 1),
    "ns2d": (2, (16, 16), (0, 0), (3, 3), (32, 32), 128, 1),
    "ns3d": (1, (8, 8, 10), (0, 0, 0), (0, 0, 3), (16, 8), 32, 1),
    "ns3d_both": (1, (6, 5, 7), (0, 0, 2), (0, 0, 2), (16, 8), 32, 1),
    "single": (1, (9, 33), (1, 0), (0, 2), (20,), 24, 1),
    "multi_out": (1, (9, 11), (0, 0), (0, 0), (8, 3, 5), 40, 3),
    "ragged_hidden": (2, (7, 37), (0, 0), (2, 1), (40, 24), 96, 2),  # three hidden chunks of 32: the fp32 kernel (hid > 64)
    "tc_two_chunks": (2, (13, 29), (0, 0), (1, 2), (40, 24), 48, 2),  # 64 channels, hid <= 64: the tcgen05 kernel, chunks 32 + 16, ragged tiles
    "tc_wide": (3, (31, 17), (2, 0), (0, 3), (64,), 64, 1),           # one 64-channel source, two full chunks
    "tc_one_chunk": (3, (29, 23), (1, 0), (0, 2), (40, 24), 20, 1),   # hid < 32, one output: the warp-specialised tcgen05 kernel
}


def grad_sample_indices(i, numel, n=64):
    """Fixed pseudo-random element indices (into the flattened real view) at which the golden fixtures keep the gradient of
    parameter number i -- element-wise model gradients without storing 16 M floats."""
    import numpy as np

    rng = np.random.default_rng(1234 + i)
    return np.sort(rng.choice(numel, size=min(n, numel), replace=False))
