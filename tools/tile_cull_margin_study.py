#!/usr/bin/env python
"""CPU study of the safety margin of the tile culling rule (common.cuh: rs_cull_limit + rs_splat_touches_rect), the test
behind the tight tile lists and the per-warp culling of the compositing kernels.

For random splats (conics from random 2-D covariances with the projection's 0.3 px^2 blur, eigenvalue ratios up to the
256:1 the rule accepts, opacities log-uniform in [1/255, 1]) against one 16x16 tile, it restates the rule in float32
(numpy, one rounding per operation as in the kernel; __fdividef as a float32 division, logf / __expf as numpy's float32
functions) and compares with what the compositing does at the 256 pixel centres of the tile, evaluated (a) in float32 in the
kernel's operation order (RasterizeToPixels3DGSFwd.cu:136-149 as raster_fwd.cu replays it) and (b) in float64.
A VIOLATION is a tile the rule culls although some pixel centre is used (alpha >= 1/255 and sigma >= 0).

    python tools/tile_cull_margin_study.py [--trials 4000000] > profiles/r02_tile_cull_margin_study.txt"""
import argparse

import numpy as np

f32 = np.float32


def fma(a, b, c):  # single rounding
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def cull_limit(a, b, c, op):
    det = a * c - b * b
    L = np.log(op * f32(255.0)).astype(f32)
    lim = (L + f32(1e-3) * (f32(1.0) + np.abs(L))).astype(f32)
    ok = (a > 0) & (c > 0) & (det > 0) & (a * c <= f32(256.0) * det) & ~np.isnan(L)
    out = np.where(ok, lim, f32(3e38)).astype(f32)
    return np.where(op < f32(1.0 / 255.0) * f32(0.999), f32(-3e38), out).astype(f32)


def touches(cx, cy, qa, qb, qc, limit, x0, x1, y0, y1):
    dx = cx - np.minimum(np.maximum(cx, x0), x1)
    dy = cy - np.minimum(np.maximum(cy, y0), y1)
    pyv = np.minimum(np.maximum(cy + (qb * dx) / qc, y0), y1)
    d2 = cy - pyv
    qv = f32(0.5) * (qa * dx * dx + qc * d2 * d2) + qb * dx * d2
    pxh = np.minimum(np.maximum(cx + (qb * dy) / qa, x0), x1)
    d1 = cx - pxh
    qh = f32(0.5) * (qa * d1 * d1 + qc * dy * dy) + qb * d1 * dy
    qmin = np.zeros_like(qv)
    qmin = np.where(dx != 0, qv, qmin)
    qmin = np.where(dy != 0, np.where(dx != 0, np.minimum(qv, qh), qh), qmin)
    return ~(qmin > limit)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=4_000_000)
    ap.add_argument("--chunk", type=int, default=100_000)
    args = ap.parse_args()
    rng = np.random.default_rng(2024)
    tot = culled = viol32 = viol64 = guarded = 0
    worst32 = worst64 = 0.0
    px = (np.arange(16, dtype=f32) + f32(0.5))
    for start in range(0, args.trials, args.chunk):
        n = min(args.chunk, args.trials - start)
        # covariance = R diag(l1, l2) R^T + 0.3 I, l in [0.01, 900] px^2 log-uniform; half of the splats strongly anisotropic
        l1 = np.exp(rng.uniform(np.log(0.01), np.log(900.0), n))
        ratio = np.where(rng.random(n) < 0.5, np.exp(rng.uniform(0, np.log(2000.0), n)), np.exp(rng.uniform(0, np.log(8.0), n)))
        l2 = np.maximum(l1 / ratio, 1e-4)
        th = rng.uniform(0, np.pi, n)
        c_, s_ = np.cos(th), np.sin(th)
        sxx = c_ * c_ * l1 + s_ * s_ * l2 + 0.3
        syy = s_ * s_ * l1 + c_ * c_ * l2 + 0.3
        sxy = c_ * s_ * (l1 - l2)
        det = sxx * syy - sxy * sxy
        qa, qb, qc = (syy / det).astype(f32), (-sxy / det).astype(f32), (sxx / det).astype(f32)
        op = np.exp(rng.uniform(np.log(1.0 / 255.0), 0.0, n)).astype(f32)
        # centres concentrated around the tile border region where the decision is made
        reach = np.sqrt(2.0 * np.maximum(np.log(255.0 * op.astype(np.float64)), 1e-3) * np.maximum(sxx, syy))
        cx = (8.0 + rng.uniform(-1, 1, n) * (8.0 + 1.3 * reach)).astype(f32)
        cy = (8.0 + rng.uniform(-1, 1, n) * (8.0 + 1.3 * reach)).astype(f32)
        limit = cull_limit(qa, qb, qc, op)
        keep = touches(cx, cy, qa, qb, qc, limit, f32(0.5), f32(15.5), f32(0.5), f32(15.5))
        guarded += int((limit > 1e38).sum())
        # compositing at the 256 pixel centres, float32 in the kernel's order
        dx = (cx[:, None, None] - px[None, None, :]).astype(f32)
        dy = (cy[:, None, None] - px[None, :, None]).astype(f32)
        A, B, C = qa[:, None, None], qb[:, None, None], qc[:, None, None]
        tc = (C * dy).astype(f32) * dy
        s = fma(dx, (A * dx).astype(f32), tc)  # [n,16,16] by broadcasting
        sigma = fma(dy, (B * dx).astype(f32), (s * f32(0.5)).astype(f32))
        alpha32 = np.minimum(f32(0.999), op[:, None, None] * np.exp(-sigma).astype(f32))
        used32 = (~((sigma < 0) | (alpha32 < f32(1.0 / 255.0)))).reshape(n, -1).any(axis=1)
        a32max = np.where(sigma < 0, 0, alpha32).reshape(n, -1).max(axis=1)
        # float64
        dx64 = cx.astype(np.float64)[:, None, None] - px.astype(np.float64)[None, None, :]
        dy64 = cy.astype(np.float64)[:, None, None] - px.astype(np.float64)[None, :, None]
        s64 = 0.5 * (qa.astype(np.float64)[:, None, None] * dx64 * dx64 + qc.astype(np.float64)[:, None, None] * dy64 * dy64) \
            + qb.astype(np.float64)[:, None, None] * dx64 * dy64
        a64max = (op.astype(np.float64)[:, None, None] * np.exp(-s64)).reshape(n, -1).max(axis=1)
        cut = ~keep
        tot += n
        culled += int(cut.sum())
        viol32 += int((cut & used32).sum())
        viol64 += int((cut & (a64max >= 1.0 / 255.0)).sum())
        if cut.any():
            worst32 = max(worst32, float(a32max[cut].max()) * 255.0)
            worst64 = max(worst64, float(a64max[cut].max()) * 255.0)
    print("# tools/tile_cull_margin_study.py -- CPU restatement (float32, numpy) of rs_cull_limit + rs_splat_touches_rect against")
    print("# the compositing's own evaluation at the 256 pixel centres of one tile")
    print(f"trials                                   {tot}")
    print(f"tiles culled by the rule                 {culled} ({culled / tot:.3f})")
    print(f"rule declined (degenerate / > 256:1)     {guarded} ({guarded / tot:.3f})  -> never culled")
    print(f"VIOLATIONS vs float32 compositing        {viol32}")
    print(f"VIOLATIONS vs float64 evaluation         {viol64}")
    print(f"largest alpha*255 inside a culled tile   float32 {worst32:.6f}   float64 {worst64:.6f}   (must stay < 1)")


if __name__ == "__main__":
    main()
