"""The algebra the B200 blend backward (csrc/backward_pipe.cu) rests on, checked on the CPU against the C oracle's
restatement of the reference (backward.cu:399-557): walking a pixel's list FRONT TO BACK with the forward's own running
(T, S) and using  dL/dalpha_i = T_i (c_i . g) - [ (C - S_i) . g + T_final (bg . g) ] / (1 - alpha_i)
gives the same nine sums per Gaussian as the reference's back-to-front walk with its "colour behind" recurrence, and the
per-Gaussian sums factor as the kernel's flush does (sum g dx, sum g dy, sum g dx dx, ... scaled once per splat)."""
import numpy as np

from fateavatar_b200 import scenes
from oracle import oracle as orc
from util import oracle_forward


def test_front_to_back_pair_gradients_equal_the_reference_recurrence():
    sc = scenes.config1_scene(P=400, W=48, H=40, seed=7)
    o = oracle_forward(orc, sc)
    W, H, P = o["W"], o["H"], o["P"]
    g_pix = np.random.default_rng(1).standard_normal((3, H, W)).astype(np.float32)
    ref = orc.backward(o, g_pix)
    bg = np.asarray(sc["bg"], np.float64)
    m2, co, rgb = o["means2D"].astype(np.float64), o["conic_opacity"].astype(np.float64), o["rgb"].astype(np.float64)
    acc = np.zeros((P, 9))  # sum g dx, g dy, g dx dx, g dx dy, g dy dy, g, w g_r, w g_g, w g_b
    gx_tiles = (W + 15) // 16
    for y in range(H):
        for x in range(W):
            lo, hi = o["ranges"][(y // 16) * gx_tiles + x // 16]
            n_c, T_final = int(o["n_contrib"][y, x]), float(o["final_T"][y, x])
            g = g_pix[:, y, x].astype(np.float64)
            C = o["color"][:, y, x].astype(np.float64) - T_final * bg  # accumulated colour without the background
            T, S = 1.0, np.zeros(3)
            for pos in range(int(hi - lo)):
                if pos >= n_c:  # the reference's backward skips positions behind the last contributor
                    break
                i = int(o["point_list"][lo + pos])
                dx, dy = m2[i, 0] - x, m2[i, 1] - y
                power = -0.5 * (co[i, 0] * dx * dx + co[i, 2] * dy * dy) - co[i, 1] * dx * dy
                if power > 0.0:
                    continue
                G = np.exp(power)
                alpha = min(0.99, co[i, 3] * G)
                if alpha < 1.0 / 255.0:
                    continue
                w = alpha * T
                S = S + w * rgb[i]
                dLda = T * (rgb[i] @ g) - ((C - S) @ g + T_final * (bg @ g)) / (1.0 - alpha)
                gg = G * dLda
                acc[i] += (gg * dx, gg * dy, gg * dx * dx, gg * dx * dy, gg * dy * dy, gg, w * g[0], w * g[1], w * g[2])
                T *= 1.0 - alpha
    # the kernel's flush (backward_pipe.cu: flush_batch): the per-splat factors are applied once per (block, splat)
    cx, cy, cz, op = co[:, 0], co[:, 1], co[:, 2], co[:, 3]
    mean2d = np.stack([-(cx * acc[:, 0] + cy * acc[:, 1]) * op * 0.5 * W, -(cz * acc[:, 1] + cy * acc[:, 0]) * op * 0.5 * H], 1)
    conic = np.stack([-0.5 * op * acc[:, 2], -0.5 * op * acc[:, 3], -0.5 * op * acc[:, 4]], 1)
    close = lambda a, b: np.abs(a - b).max() <= 2e-4 * max(np.abs(b).max(), 1e-12)
    assert close(mean2d, ref["dL_dmeans2D"][:, :2])
    assert close(acc[:, 5], ref["dL_dopacity"].reshape(-1))
    assert close(acc[:, 6:9], ref["dL_dcolors"])
    rc = np.asarray(ref["dL_dconic"], np.float64).reshape(P, -1)
    assert close(conic, rc[:, [0, 1, rc.shape[1] - 1]])  # (x, y, w) of the reference's float4 / (x, y, z) of a float3
