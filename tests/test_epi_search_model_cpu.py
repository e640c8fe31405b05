"""Model of the warp-shared epipolar search of the depth-filter seed kernels (pl-svo_b200/csrc/align2d_kernel.cu,
`warp_epipolar_search`) against the sequential loop it replaces (src/matcher.cpp:354-391):

    for i in range(n_steps):  uv += step  (rounded double additions)
        pxi = round(project(uv));  if pxi == last_checked: continue;  last_checked = pxi
        if not in_frame(pxi): continue
        if score(pxi) < best: best, uv_best = score, uv

The kernel walks the first `serial_steps` steps on the seed's own thread and hands the rest to the 32 lanes of the warp,
32 consecutive steps per pass: lane j reaches its uv by j rounded additions from the pass's first value, the
"same pixel as the previous step" test compares with the neighbouring lane (lane 0 with the last pixel of the previous
pass), and the pass winner is the lowest score, ties to the lowest lane.  This file states that schedule in NumPy and
checks it against the loop on adversarial inputs (repeated pixels, ties, out-of-frame stretches, every split point)."""
import numpy as np
import pytest


def sequential(uv0, step, n, project, in_frame, score):
    uv = np.float64(uv0)
    last = (0, 0)  # `last_checked_pxi` starts at (0, 0)
    best, uv_best = 2000 * 64, np.float64(0.0)
    for _ in range(n):
        pxi = project(uv)
        if pxi != last:
            last = pxi
            if in_frame(pxi):
                z = score(pxi)
                if z < best:
                    best, uv_best = z, uv
        uv = np.float64(uv + step)
    return best, uv_best


def hybrid(uv0, step, n, project, in_frame, score, serial_steps, lanes=32):
    # own thread: steps [0, min(n, serial_steps))
    uv = np.float64(uv0)
    last = (0, 0)
    best, uv_best = 2000 * 64, np.float64(0.0)
    k_done = 0
    while k_done < min(n, serial_steps):
        pxi = project(uv)
        if pxi != last:
            last = pxi
            if in_frame(pxi):
                z = score(pxi)
                if z < best:
                    best, uv_best = z, uv
        uv = np.float64(uv + step)
        k_done += 1
    # whole warp: passes of `lanes` consecutive steps
    base, prev = uv, last
    k0 = k_done
    while k0 < n:
        u = []
        for lane in range(lanes):  # lane j: j rounded additions from the pass's first value
            x = base
            for _ in range(lane):
                x = np.float64(x + step)
            u.append(x)
        base = np.float64(u[lanes - 1] + step)
        pxis = [project(x) for x in u]
        neighbour = [prev] + pxis[:-1]
        prev = pxis[lanes - 1]
        zs = []
        for lane in range(lanes):
            ok = (k0 + lane < n) and pxis[lane] != neighbour[lane] and in_frame(pxis[lane])
            zs.append(score(pxis[lane]) if ok else 0x7FFFFFFF)
        zmin = min(zs)
        lmin = zs.index(zmin)  # ties to the lowest lane
        if zmin < best:
            best, uv_best = zmin, u[lmin]
        k0 += lanes
    return best, uv_best


@pytest.mark.parametrize("seed", range(12))
def test_hybrid_schedule_equals_the_sequential_loop(seed):
    rng = np.random.default_rng(9000 + seed)
    for trial in range(40):
        n = int(rng.choice([1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 257, int(rng.integers(1, 400))]))
        uv0 = np.float64(rng.uniform(-0.4, 0.4))
        # steps from far below a pixel (long runs of repeated pixels) to several pixels, irrational-looking so that the
        # repeated additions round differently from a multiplication
        step = np.float64(rng.choice([1e-4, 7e-4, 3.3e-3, 0.011]) * rng.uniform(0.5, 1.5) * rng.choice([-1, 1]))
        fx, cx, scale = 315.5, 376.0, float(rng.choice([1, 2, 4]))
        lo, hi = sorted(rng.integers(0, 760, 2).tolist())
        table = rng.integers(0, 6, 4096)  # few distinct scores: many ties
        bias = int(rng.integers(0, 2000 * 64 + 50))  # sometimes no score beats the initial best

        def project(uv):
            return (int(np.floor((fx * uv + cx) / scale + 0.5)), 7)

        def in_frame(pxi):
            return lo <= pxi[0] <= hi

        def score(pxi):
            return int(table[pxi[0] % 4096]) + bias

        ref = sequential(uv0, step, n, project, in_frame, score)
        for serial_steps in (0, 16, 32, n, n + 5):
            got = hybrid(uv0, step, n, project, in_frame, score, serial_steps)
            assert got[0] == ref[0], (seed, trial, n, serial_steps)
            assert got[1].tobytes() == ref[1].tobytes(), (seed, trial, n, serial_steps)


def test_repeated_additions_are_not_a_multiplication():
    """Why lane j applies j additions instead of computing uv0 + j * step: the two differ in the last bits."""
    uv0, step = np.float64(0.1234567), np.float64(0.0007654321)
    x = uv0
    differs = False
    for j in range(1, 200):
        x = np.float64(x + step)
        differs |= x != np.float64(uv0 + j * step)
    assert differs
