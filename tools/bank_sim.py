"""Shared-memory bank-conflict model of the forward gather (developer tool, CPU only).

For the C2 geometry (128x128 image, P = 184) it replays which 16-byte chunks the 8 lanes of
a quarter-warp read in one LDS.128 and counts wavefronts = max number of distinct addresses
per bank group.  Results quoted in DESIGN.md:
  images per pixel record 4 / 8 / 16 / 32  ->  1.57 / 1.34 / 1.16 / 1.00 wavefronts per quarter
(ncu measured 1.73 and 1.165 for 4 and 16).  `python tools/bank_sim.py`
"""
import numpy as np

X, P, PAD = 128, 184, 28


def coords(theta, j, i):
    ang = np.float32(-theta)
    c, s = np.cos(ang), np.sin(ang)
    wm1 = np.float32(P - 1)
    xoff = (wm1 - (c * wm1 - s * wm1)) / 2
    yoff = (wm1 - (s * wm1 + c * wm1)) / 2
    return c * j - s * i + xoff, s * j + c * i + yoff


def multiplier(depth_log2, thetas):
    """quarter-warp = (8 >> a) adjacent rays x (1 << a) image groups, rows synchronised."""
    nr = slots = 8 >> depth_log2
    tot = cnt = 0
    for th in thetas:
        cc = np.cos(np.float32(-th))
        for j0 in range(32, 152, nr):
            j = np.arange(j0, j0 + nr, dtype=np.float32)
            _, yy = coords(th, j, np.float32(0))
            for r in range(PAD + 10, PAD + 110, 3):
                i = np.ceil((r - yy) / cc)
                for n in range(2):
                    x, y = coords(th, j, (i + n).astype(np.float32))
                    col, row = np.floor(x).astype(int), np.floor(y).astype(int)
                    wf = max(len(set(zip(row[col % slots == sl], col[col % slots == sl]))) for sl in range(slots))
                    tot += wf
                    cnt += 1
    return tot / cnt


if __name__ == "__main__":
    thetas = np.concatenate([np.linspace(0, np.pi / 4, 46), np.linspace(3 * np.pi / 4, np.pi, 45, endpoint=False)])
    for a in range(4):
        print(f"{4 << a:2d} images per record ({8 >> a} rays x {1 << a} groups per quarter-warp): "
              f"{multiplier(a, thetas):.3f} wavefronts per quarter-warp load")
