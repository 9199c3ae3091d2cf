"""Probe (run on the GPU box): which leading / stride byte-offset convention makes a [rows][features] fp16 tile in the
no-swizzle core-matrix layout readable K-major and MN-major by tcgen05.mma.kind::f16.  Prints one line per variant."""
import itertools
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from minimarl_b200 import _lib  # noqa: E402


def tile_image(mat):
    """[rows][feats] -> 64 KB image, chunk-major: off(r, f) = (f//8)*CH + (r//8)*128 + (r%8)*16 + (f%8)*2, CH = rows/8*128."""
    rows, feats = mat.shape
    ch = rows // 8 * 128
    img = np.zeros(32768, dtype=np.float16)
    r, f = np.meshgrid(np.arange(rows), np.arange(feats), indexing="ij")
    off = (f // 8) * ch + (r // 8) * 128 + (r % 8) * 16 + (f % 8) * 2
    img[off // 2] = mat.astype(np.float16)
    return img, ch


def run(imgA, imgB, M, N, steps, a, b, ta, tb):
    dev = torch.device("cuda:0")
    A = torch.from_numpy(imgA.view(np.uint8)).to(dev)
    B = torch.from_numpy(imgB.view(np.uint8)).to(dev)
    D = torch.zeros((128, N), device=dev)
    _lib.check(_lib.lib().mm_tc_raw_f16_selftest(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), M, N, steps, *a, *b, ta, tb, _lib.stream()))
    torch.cuda.synchronize()
    return D.cpu().numpy()


def main():
    rng = np.random.default_rng(0)
    out = {}
    # 1. both K-major: D[r][n] = sum_f X[r][f] W[n][f]
    X = rng.integers(-4, 5, (128, 32)).astype(np.float32)
    W = rng.integers(-4, 5, (32, 32)).astype(np.float32)
    ix, chx = tile_image(X)
    iw, chw = tile_image(W)
    D = run(ix, iw, 128, 32, 2, (2 * chx, chx, 128), (2 * chw, chw, 128), 0, 0)
    out["kk"] = bool(np.array_equal(D, X @ W.T))
    # 2. both MN-major: D[g][x] = sum_r G[r][g] X[r][x]
    G = rng.integers(-4, 5, (128, 128)).astype(np.float32)
    ig, chg = tile_image(G)
    want = G.T @ X
    for name, a, b in (("mnmn_doc", (256, 128, chg), (256, 128, chx)), ("mnmn_swapped", (256, chg, 128), (256, chx, 128))):
        D = run(ig, ix, 128, 32, 8, a, b, 1, 1)
        out[name] = bool(np.array_equal(D, want))
    # 3. A K-major (rows x 96 gate features), B MN-major (weight stored [gate][in], read as [in][gate])
    Wg = rng.integers(-4, 5, (96, 32)).astype(np.float32)
    iwg, chwg = tile_image(Wg)
    want = G[:, :96] @ Wg
    for name, b in (("k_mn_doc", (256, 128, chwg)), ("k_mn_swapped", (256, chwg, 128))):
        D = run(ig, iwg, 128, 32, 6, (2 * chg, chg, 128), b, 0, 1)
        out[name] = bool(np.array_equal(D, want))
    # 4. MN-major A with M = 64 (lane placement of a 64-row accumulator), N = 48
    X48 = rng.integers(-4, 5, (128, 48)).astype(np.float32)
    ix48, chx48 = tile_image(X48)
    want = G[:, :64].T @ X48
    D = run(ig, ix48, 64, 48, 8, (256, 128, chg), (256, 128, chx48), 1, 1)
    lanes = np.concatenate([np.arange(16) + 32 * q for q in range(4)])
    out["mn_m64_lanes_0_15_of_each_quarter"] = bool(np.array_equal(D[lanes], want))
    # 5. N = 80 and N = 96 accumulators (M = 128, MN-major both)
    X96 = rng.integers(-4, 5, (128, 96)).astype(np.float32)
    ix96, chx96 = tile_image(X96)
    for n in (80, 96):
        D = run(ig, ix96, 128, n, 8, (256, 128, chg), (256, 128, chx96), 1, 1)
        out[f"mnmn_n{n}"] = bool(np.array_equal(D, (G.T @ X96)[:, :n]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
