"""Coverage check (numpy, CPU) of tc_rowgemm.cuh's epilogue column split: for G warp groups taking J 16-column chunks per round
(default G=2, J=2; opt-in G=4, J=1 and its prefetching twin), every element of a 128-row x N_t-column tile below column N is
owned by exactly one (warp, lane, register) -- the tcgen05.ld 16x256b ownership the kernel relies on."""
import numpy as np


def cover(G, J, N_t, N, n_base=0, prefetch=False):
    cnt = np.zeros((128, N_t), int)
    for grp in range(G):
        for q in range(4):
            for lane in range(32):
                rows = [q * 32 + (hr >> 1) * 16 + (hr & 1) * 8 + (lane >> 2) for hr in range(4)]   # index hh*2 + rr
                if prefetch:
                    ci = grp
                    rounds = []
                    while ci * 16 < N_t and n_base + ci * 16 < N:
                        rounds.append([ci * 16])
                        ci += G
                else:
                    rounds = []
                    ci = grp
                    while ci * 16 < N_t:
                        c0a, c0b = ci * 16, (ci + G) * 16
                        if n_base + c0a >= N:
                            break
                        chunk = [c0a]
                        if J == 2 and c0b < N_t and n_base + c0b < N:
                            chunk.append(c0b)
                        rounds.append(chunk)
                        ci += J * G
                for chunk in rounds:
                    for c0 in chunk:
                        for hh in range(2):
                            for rep in range(2):
                                for rr in range(2):
                                    c = c0 + rep * 8 + 2 * (lane & 3)
                                    for d in range(2):
                                        if n_base + c + d < N:
                                            cnt[rows[hh * 2 + rr], c + d] += 1
    want = np.zeros((128, N_t), int)
    want[:, : max(0, min(N_t, N - n_base))] = 1
    assert (cnt == want).all(), (G, J, N_t, N, n_base, np.argwhere(cnt != want)[:4])


for N_t, N, n_base in [(256, 481, 0), (256, 481, 256), (128, 120, 0), (64, 63, 0), (48, 33, 0), (16, 9, 0), (240, 240, 0), (32, 31, 0)]:
    cover(2, 2, N_t, N, n_base)
    cover(4, 1, N_t, N, n_base)
    cover(4, 1, N_t, N, n_base, prefetch=True)
    cover(2, 1, N_t, N, n_base)          # the analysis kernel's 16-loader-warp instantiation calls the two groups in turn
print("epilogue column split: every element owned exactly once")
