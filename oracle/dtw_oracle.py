"""CPU ORACLE (numpy) for the evaluation metrics of SURVEY.md §8(f)-4.  TEST INFRASTRUCTURE, not product code.

The reference calls the third-party package `dtw_c` (`dtw.calc_mcd`, `dtw.dtw_org_to_trg`; call sites
train_gru_cyclevae_gauss_batch.py:679-688,1435-1439 and decode_gru-cyclevae_gauss.py:334-393).  dtw_c is NOT in the
reference tree and is unpinned in tools/requirements.txt, so its source could not be read: PARITY UNPINNED for this row.
What is restated here is the published definition the call sites rely on:

    MCD(x, y)  = (10 / ln 10) * sqrt(2 * sum_d (x_d - y_d)^2)   [dB] per frame, mean over frames
    DTW        : G(i,j) = d(i,j) + min(G(i-1,j-1), G(i-1,j), G(i,j-1)), G(0,0) = d(0,0), end point (N-1, M-1),
                 ties resolved diagonal first, then (i-1,j), then (i,j-1); org_to_trg keeps, for every target frame,
                 the LAST source frame the path pairs with it.
"""
import numpy as np

MCD_K = (10.0 / np.log(10.0)) * np.sqrt(2.0)


def mcd_frames(x, y):
    d = np.asarray(x, dtype=np.float64) - np.asarray(y, dtype=np.float64)
    return MCD_K * np.sqrt(np.sum(d * d, axis=1))


def calc_mcd(x, y):
    m = mcd_frames(x, y)
    return float(m.mean()), float(m.std())


def dtw_org_to_trg(org, trg, dist=None):
    """-> (aligned_org [M, D], path [M], mean MCD over target frames, path steps, accumulated cost)."""
    org, trg = np.asarray(org, dtype=np.float64), np.asarray(trg, dtype=np.float64)
    N, M = len(org), len(trg)
    if dist is None:
        dist = MCD_K * np.sqrt(np.maximum(((org[:, None, :] - trg[None, :, :]) ** 2).sum(-1), 0.0))
    dist = np.asarray(dist, dtype=np.float32)
    G = np.full((N, M), np.inf, dtype=np.float32)
    D = np.zeros((N, M), dtype=np.uint8)
    for i in range(N):
        for j in range(M):
            if i == 0 and j == 0:
                best, d = np.float32(0), 0
            else:
                best, d = np.float32(np.inf), 0
                if i > 0 and j > 0:
                    best = G[i - 1, j - 1]
                if i > 0 and G[i - 1, j] < best:
                    best, d = G[i - 1, j], 1
                if j > 0 and G[i, j - 1] < best:
                    best, d = G[i, j - 1], 2
            G[i, j] = np.float32(dist[i, j] + best)
            D[i, j] = d
    path = -np.ones(M, dtype=np.int32)
    i, j, steps = N - 1, M - 1, 0
    while True:
        if path[j] < 0:
            path[j] = i
        steps += 1
        if i == 0 and j == 0:
            break
        d = D[i, j]
        if d == 0:
            i, j = i - 1, j - 1
        elif d == 1:
            i -= 1
        else:
            j -= 1
    mean = float(np.mean(dist[path, np.arange(M)].astype(np.float64)))
    return org[path], path, mean, steps, float(G[N - 1, M - 1])
