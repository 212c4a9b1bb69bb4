"""Index-math prototype of the kw-tap-folded small-N convolution planned for round 2 — CPU only.

Why: a 3x3 conv with cout <= 64 issues one M128 x N(cout) x K16 MMA per (tap, 16 input channels); every MMA re-reads its
4 KB A slice from shared memory whatever N is, so the tensor pipe idles behind the A operand (40-80 cycles per MMA against
16-32 of tensor work; profiles/r1m/timeline_light_dense3x3.txt).  Folding the three taps of a filter ROW into N makes the
same A slice feed three times the columns:

    D[j][(s, co)] = sum_{r, ci}  X[h + r - 1][w0 - 1 + j][ci] * W[co][ci][r][s]          j = 0..127 (TMEM lane)
    out[h][w0 + i][co] = D[i][(0, co)] + D[i + 1][(1, co)] + D[i + 2][(2, co)]           i = 0..125

  * A: the halo box is exactly {Ck, 128, 1, rows + 2, 1} starting at column w0 - 1 (no extra halo columns); the three filter
    rows are three row-shifted views of it, as today.  K walk = (r, channel chunk): a third of today's MMAs.
  * B: weights packed as Bf[(s, co)][(r, ci)] = W[co][ci][r][s]  (N = 3 * cout rows, K = 3 * cin), K-major.
  * Tiles advance by 126 output pixels per 128 lanes (w0 = 126 * t): lanes 126/127 only feed their neighbours.
  * Epilogue: lane i needs columns (1, co) of lane i + 1 and (2, co) of lane i + 2: two `shfl.down` inside a warp plus the
    first two lanes of the next warp's TMEM quarter through a 2 x cout fp32 shared-memory patch (the only cross-warp traffic);
    then scale/shift/activation/residual as today; 126-pixel rows are stored with per-lane 16-byte stores (a fixed TMA box
    cannot express 32/32/32/30).

`python tools/proto_kwfold.py` checks the algebra (including image borders, ragged last tiles and the 126-pixel tiling)
against torch.nn.functional.conv2d in fp64.
"""
import torch
import torch.nn.functional as F

LANES = 128
OUT_PER_TILE = LANES - 2


def pack_kwfold(w):
    """W[co][ci][3][3] -> Bf[(s, co)][(r, ci)]: rows = 3*cout (tap column s major), K = 3*cin (tap row r major)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    return w.permute(3, 0, 2, 1).reshape(kw * co, kh * ci).contiguous()


def conv_kwfold(x, w):
    """x: [n, h, wd, ci] (NHWC), w: [co, ci, 3, 3] -> [n, h, wd, co], computed the way the planned kernel would."""
    n, h, wd, ci = x.shape
    co = w.shape[0]
    bf = pack_kwfold(w)                                            # [(s, co), (r, ci)]
    xp = F.pad(x, (0, 0, 1, LANES, 1, 1))                           # TMA zero fill: one column left, a tile's worth right, one row each side
    out = torch.zeros(n, h, wd, co, dtype=x.dtype)
    for img in range(n):
        for row in range(h):
            for t in range((wd + OUT_PER_TILE - 1) // OUT_PER_TILE):
                w0 = t * OUT_PER_TILE
                # A operand: lanes j <-> input column w0 - 1 + j (padded index w0 + j), K = (r, ci) from three row-shifted views
                a = torch.cat([xp[img, row + r, w0:w0 + LANES, :] for r in range(3)], dim=1)      # [128, 3*ci]
                d = a @ bf.t()                                                                      # [128, 3*co]  (the MMAs)
                d = d.view(LANES, 3, co)
                # epilogue: out_i = D[i][0] + D[i+1][1] + D[i+2][2]
                o = d[0:OUT_PER_TILE, 0] + d[1:OUT_PER_TILE + 1, 1] + d[2:OUT_PER_TILE + 2, 2]
                valid = min(OUT_PER_TILE, wd - w0)
                out[img, row, w0:w0 + valid] = o[:valid]
    return out


def main():
    torch.manual_seed(0)
    for (n, h, wd, ci, co) in [(1, 5, 126, 16, 32), (2, 4, 300, 32, 32), (1, 3, 7, 16, 48), (1, 6, 252, 64, 64), (1, 2, 127, 32, 16)]:
        x = torch.randn(n, h, wd, ci, dtype=torch.float64)
        w = torch.randn(co, ci, 3, 3, dtype=torch.float64)
        ref = F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
        got = conv_kwfold(x, w)
        err = (got - ref).abs().max().item()
        assert err < 1e-10, (n, h, wd, ci, co, err)
        mmas_now = 9 * (ci // 16)
        mmas_fold = 3 * (ci // 16)
        print(f"n={n} h={h} w={wd} cin={ci} cout={co}: max err {err:.1e}; MMAs per 128-lane tile {mmas_now} (N={co}) -> {mmas_fold} (N={3 * co}), "
              f"pixel efficiency {OUT_PER_TILE}/{LANES}")
    print("ok")


if __name__ == "__main__":
    main()
