#!/usr/bin/env python
"""Which float32 association do the torch CUDA ops inside the reference's apply_transform() (main.py:183-228) use?

The fused rigid transform must reproduce them bit for bit, otherwise a moved mean differs by an ulp and a ceil()/floor()
in the projection / tile count can flip.  For each op this prints the fraction of outputs that are bit-identical to
candidate evaluation orders (emulated in float64 -> float32 on the CPU; products of two floats are exact in float64):
  * torch.linalg.norm(q) on a [4] tensor                          (main.py:207)
  * torch.matmul(means - center, rot_mat.T), [n,3] x [3,3]        (main.py:213)
Writes one JSON object to stdout (kept under profiles/)."""
import json
import sys

import numpy as np
import torch

dev = "cuda:0"
f32 = np.float32


def r32(x):
    return np.asarray(x, np.float64).astype(np.float32)


def fma(a, b, c):  # round32(a*b + c) with an exact product (double rounding is astronomically rare)
    return r32(a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64))


def mul(a, b):
    return r32(a.astype(np.float64) * b.astype(np.float64))


def add(a, b):
    return r32(a.astype(np.float64) + b.astype(np.float64))


def main():
    out = {}
    rng = np.random.default_rng(0)
    # ---- norm of a 4-vector -------------------------------------------------------------------------------------------
    Q = rng.standard_normal((512, 4)).astype(f32)
    got = np.array([float(torch.linalg.norm(torch.from_numpy(q).to(dev)).cpu()) for q in Q], f32)
    got_batched = torch.linalg.norm(torch.from_numpy(Q).to(dev), dim=-1).cpu().numpy()
    w, x, y, z = Q.T
    sq = lambda v: mul(v, v)
    cands = {
        "sqrt(((w2+x2)+y2)+z2) rounded squares": np.sqrt(add(add(add(sq(w), sq(x)), sq(y)), sq(z))),
        "sqrt((w2+x2)+(y2+z2)) rounded squares": np.sqrt(add(add(sq(w), sq(x)), add(sq(y), sq(z)))),
        "sqrt(fma(z,z,fma(y,y,fma(x,x,w*w))))": np.sqrt(fma(z, z, fma(y, y, fma(x, x, sq(w))))),
        "sqrt(fma(w,w,0)+..) pairwise fma: fma(y,y,w2f)+fma(z,z,x2f)": np.sqrt(add(fma(y, y, sq(w)), fma(z, z, sq(x)))),
        "exactly rounded (float64 sum)": r32(np.sqrt((Q.astype(np.float64) ** 2).sum(1))),
    }
    # exhaustive family: every binary summation tree over the four squares, every addition with a single-leaf operand
    # optionally contracted to fma(x, x, other)
    import itertools

    def trees(idx):  # yields (label, value) for all ways to sum the squares of the columns in idx
        if len(idx) == 1:
            yield f"s{idx[0]}", sq(Q[:, idx[0]])
            return
        seen = set()
        for r in range(1, len(idx)):
            for left in itertools.combinations(idx, r):
                right = tuple(i for i in idx if i not in left)
                if (right, left) in seen:
                    continue
                seen.add((left, right))
                for (la, va), (lb, vb) in itertools.product(trees(left), trees(right)):
                    yield f"({la}+{lb})", add(va, vb)
                    if len(right) == 1:
                        yield f"fma(x{right[0]},x{right[0]},{la})", fma(Q[:, right[0]], Q[:, right[0]], va)
                    if len(left) == 1:
                        yield f"fma(x{left[0]},x{left[0]},{lb})", fma(Q[:, left[0]], Q[:, left[0]], vb)

    family = {}
    for label, v in trees((0, 1, 2, 3)):
        family.setdefault(label, float((np.sqrt(v).astype(f32) == got).mean()))
    best = sorted(family.items(), key=lambda kv: -kv[1])[:8]
    out["linalg_norm_vec4_best_of_family"] = {"n_candidates": len(family), "best": best}
    out["linalg_norm_vec4_equals_batched_dim-1"] = float((got == got_batched).mean())
    # x / norm(x) per body vs batched (what a host-side normalisation of the [K,4] pose table would compute)
    qn_single = np.stack([(torch.from_numpy(q).to(dev) / torch.linalg.norm(torch.from_numpy(q).to(dev))).cpu().numpy() for q in Q[:128]])
    tq = torch.from_numpy(Q[:128]).to(dev)
    qn_batched = (tq / torch.linalg.norm(tq, dim=-1, keepdim=True)).cpu().numpy()
    out["normalised_quat_single_equals_batched"] = float((qn_single == qn_batched).all(-1).mean())
    out["linalg_norm_vec4"] = {k: float((v.astype(f32) == got).mean()) for k, v in cands.items()}
    out["linalg_norm_batched_dim-1"] = {k: float((v.astype(f32) == got_batched).mean()) for k, v in cands.items()}
    # ---- [n,3] x [3,3] matmul ---------------------------------------------------------------------------------------
    for n in (1, 7, 1000, 50_000, 1_000_000):
        A = (rng.standard_normal((n, 3)) * 0.5).astype(f32)
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        ww, xx, yy, zz = q
        R = np.array([[1 - 2 * (yy * yy + zz * zz), 2 * (xx * yy - ww * zz), 2 * (xx * zz + ww * yy)],
                      [2 * (xx * yy + ww * zz), 1 - 2 * (xx * xx + zz * zz), 2 * (yy * zz - ww * xx)],
                      [2 * (xx * zz - ww * yy), 2 * (yy * zz + ww * xx), 1 - 2 * (xx * xx + yy * yy)]]).astype(f32)
        tA, tR = torch.from_numpy(A).to(dev), torch.from_numpy(R).to(dev)
        got = torch.matmul(tA, tR.T).cpu().numpy()
        a0, a1, a2 = A[:, 0:1], A[:, 1:2], A[:, 2:3]
        b0, b1, b2 = R[:, 0][None], R[:, 1][None], R[:, 2][None]  # out[:, j] = sum_k A[:, k] * R[j, k]
        cands = {
            "fma(a2,b2,fma(a1,b1,a0*b0)) (k ascending)": fma(a2, b2, fma(a1, b1, mul(a0, b0))),
            "fma(a0,b0,fma(a1,b1,a2*b2)) (k descending)": fma(a0, b0, fma(a1, b1, mul(a2, b2))),
            "(a0*b0+a1*b1)+a2*b2 no fma": add(add(mul(a0, b0), mul(a1, b1)), mul(a2, b2)),
            "exactly rounded (float64)": r32(A.astype(np.float64) @ R.astype(np.float64).T),
        }
        out[f"matmul_n{n}"] = {k: float((v.astype(f32) == got).mean()) for k, v in cands.items()}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
