#!/usr/bin/env python
"""tools/plan_model.py -- the structural efficiency of the engine's work decomposition, computed on the CPU with the
planner's own rule (csrc/b200_launch.cuh: plan_zchunks): per test, precision and grid
    fill   = items / (rounds x 148 CTAs)              (tail of the last round)
    zwork  = len / (len + WARM)                       (warm-up planes of every z-chunk: loaded and summed, not emitted)
    yfill  = interior rows / (y-tiles x TY)           (rows of the last y-tile that fall outside the interior)
Their product is the ceiling the decomposition alone puts on the roofline fraction of an HBM-bound stencil: 0.82-0.92
at 512x256x256 (few tiles, hence short z-chunks), 0.94-0.98 at 1024x1024x512.  DESIGN.md section 8 uses it."""

CAP = 148


def plan(tiles_xy, nz, warm, cap=CAP):
    best, best_score = 1, -1.0
    for nzc in range(1, nz + 1):
        ln = (nz + nzc - 1) // nzc
        if (nz + ln - 1) // ln != nzc:
            continue
        items = tiles_xy * nzc
        rounds = (items + cap - 1) // cap
        score = items / (rounds * cap) * ln / (ln + warm)
        if score > best_score + 1e-9:
            best_score, best = score, nzc
        if ln <= 4:
            break
    ln = (nz + best - 1) // best
    items = tiles_xy * best
    rounds = (items + cap - 1) // cap
    return best, ln, items, rounds, items / (rounds * cap), ln / (ln + warm)


# test: (TY double, TY float, WARM, stencil radius in y/z)  -- b200_ops3d.cuh
OPS = {"laplacian": (24, 48, 2, 1), "wave13pt": (12, 24, 4, 2), "divergence": (12, 24, 2, 1), "gradient": (12, 24, 2, 1),
       "uxx1": (6, 12, 3, 2), "lapgsrb": (12, 24, 4, 2), "tricubic": (8, 16, 3, 1)}

if __name__ == "__main__":
    print(f"{'test':10s} {'grid':3s} {'T':1s} {'TY':>2s} {'tiles':>5s} {'nzc':>3s} {'len':>3s} {'items':>5s} {'fill':>5s} {'zwork':>5s} {'yfill':>5s} {'total':>5s}")
    for name, (tyd, tyf, warm, r) in OPS.items():
        for label, (nx, ny, ns) in (("C1", (512, 256, 256)), ("C2", (1024, 1024, 512))):
            for real, ty in (("d", tyd), ("f", tyf)):
                ntx = (nx + 127) // 128
                ylen = ny - 2 * r
                nty = (ylen + ty - 1) // ty
                nzc, ln, items, rounds, fill, work = plan(ntx * nty, ns - 2 * r, warm)
                yfill = ylen / (nty * ty)
                print(f"{name:10s} {label:3s} {real:1s} {ty:2d} {ntx * nty:5d} {nzc:3d} {ln:3d} {items:5d} {fill:5.3f} {work:5.3f} {yfill:5.3f} {fill * work * yfill:5.3f}")
