"""np_oracle.py — an INDEPENDENT NumPy restatement of plsvo::SparseImgAlign::run and
plsvo::pose_optimizer::optimizeGaussNewton, written from the reference text separately from
oracle/plsvo_oracle.cpp (rotation matrices + closed-form exp instead of quaternions,
numpy.linalg instead of a hand LDLT, vectorised patches instead of pointer walks).

TEST INFRASTRUCTURE ONLY.  Its job is to pin the C++ oracle: two independent transcriptions of
the same source must produce the same per-iteration H, Jres, chi2, n_meas and step
(tests/test_oracle_cpu.py::test_cpp_and_numpy_restatements_agree_*).  Float32 steps that the reference does in float are done with
numpy float32 scalars/arrays (IEEE, no FMA); sums over pixels run in float64 in a different
order than the reference, so agreement is to ~1e-6 relative on chi2 and ~1e-9 on H.

Reference: src/sparse_img_align.cpp (whole file), src/pose_optimizer.cpp:38-260,
src/feature.cpp:160-218, include/plsvo/frame.h:131-160; vikit NLLSSolver / robust_cost,
Sophus SE3 (un-vendored).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------
# SE3 with rotation matrices
# ------------------------------------------------------------------------------------------------
def quat_to_R(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def pose7_to_Rt(p):
    return quat_to_R(np.asarray(p[:4], float)), np.asarray(p[4:], float)


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def se3_exp(xi):
    ups, om = xi[:3], xi[3:]
    th = np.linalg.norm(om)
    W = hat(om)
    if th < 1e-10:
        R = np.eye(3) + W
        V = np.eye(3)
    else:
        R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th**2 * W @ W
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * W + (th - np.sin(th)) / th**3 * W @ W
    return R, V @ ups


def jac_xyz2uv(p):
    """Frame::jacobian_xyz2uv for an array of points [n,3] -> [n,2,6] (frame.h:138-160)."""
    x, y, zi = p[:, 0], p[:, 1], 1.0 / p[:, 2]
    zi2 = zi * zi
    J = np.zeros((len(p), 2, 6))
    J[:, 0, 0] = -zi
    J[:, 0, 2] = x * zi2
    J[:, 0, 3] = y * J[:, 0, 2]
    J[:, 0, 4] = -(1.0 + x * J[:, 0, 2])
    J[:, 0, 5] = y * zi
    J[:, 1, 1] = -zi
    J[:, 1, 2] = y * zi2
    J[:, 1, 3] = 1.0 + y * J[:, 1, 2]
    J[:, 1, 4] = -J[:, 0, 3]
    J[:, 1, 5] = -x * zi
    return J


# ------------------------------------------------------------------------------------------------
# patches
# ------------------------------------------------------------------------------------------------
def patch_geometry(u, v):
    """Patch::setPosition / computeInterpWeights (feature.cpp:189-208) for arrays of centres."""
    uf, vf = u.astype(f32), v.astype(f32)
    ui, vi = np.floor(uf).astype(np.int64), np.floor(vf).astype(np.int64)
    su = (uf - ui.astype(f32)).astype(np.float64)
    sv = (vf - vi.astype(f32)).astype(np.float64)
    w = np.stack([(1 - su) * (1 - sv), su * (1 - sv), (1 - su) * sv, su * sv], -1).astype(f32)
    return ui, vi, w


def interp_block(img, ui, vi, w, dy, dx):
    """Interpolated 4x4 block whose pixel (y,x) is the bilinear sample at (vi-2+y+dy, ui-2+x+dx)."""
    n = len(ui)
    yy = (vi[:, None, None] - 2 + dy + np.arange(4)[None, :, None])
    xx = (ui[:, None, None] - 2 + dx + np.arange(4)[None, None, :])
    I = lambda a, b: img[yy + a, xx + b].astype(f32)
    wTL, wTR, wBL, wBR = (w[:, k][:, None, None] for k in range(4))
    return ((wTL * I(0, 0) + wTR * I(0, 1)) + wBL * I(1, 0)) + wBR * I(1, 1)  # float32, reference order


def in_frame(ui, vi, cols, rows, b):
    return ~((ui < b) | (vi < b) | (ui >= cols - b) | (vi >= rows - b))


def setup_sampling(spx, epx, length):
    dif = epx - spx
    t = min(abs(dif[0]), abs(dif[1])) / max(abs(dif[0]), abs(dif[1]))
    s = t / np.sqrt(1.0 + t * t)
    corr = 2.0 * np.sqrt(1.0 + s * s)
    return int(max(1.0, length / (2.0 * 4 * corr))), dif


# ------------------------------------------------------------------------------------------------
# SparseImgAlign::run for one pair, returning the per-iteration trace
# ------------------------------------------------------------------------------------------------
def align_pair(d, b, max_level=4, min_level=2, n_iter=30, eps=1e-6):
    cam = d.cam
    R_ref, t_ref = pose7_to_Rt(d.T_ref_w[b])
    R_cur, t_cur = pose7_to_Rt(d.T_cur_w[b])
    ref_pos = -R_ref.T @ t_ref
    R = R_cur @ R_ref.T  # T_cur_from_ref
    t = t_cur - R @ t_ref
    n_pts, n_segs = d.n_pts, d.n_segs
    pt_xyz = d.pt_f[b] * np.linalg.norm(d.pt_pos[b] - ref_pos, axis=1, keepdims=True) if n_pts else np.zeros((0, 3))
    seg_alive = np.ones(n_segs, bool)
    pt_visible = np.zeros(n_pts, bool)
    trace = []
    chi2_prev, stop = 1e10, False
    n_meas = 0
    for level in range(max_level, min_level - 1, -1):
        scale = f32(1.0) / f32(1 << level)
        cols, rows = cam.width >> level, cam.height >> level
        ref_img, cur_img = d.ref_pyr[level][b], d.cur_pyr[level][b]
        cJ = abs(cam.fx) / (1 << level)
        # ---- precompute (sparse_img_align.cpp:195-378) ----
        patches_xyz, patches_px, owner = [], [], []  # owner: -1 for points, segment index otherwise
        if n_pts:
            u = d.pt_px[b][:, 0] * float(scale)
            v = d.pt_px[b][:, 1] * float(scale)
            ui, vi, _ = patch_geometry(u, v)
            vis = in_frame(ui, vi, cols, rows, 3)
            pt_visible |= vis
            pt_fresh = vis
        seg_samples = {}
        for j in range(n_segs):
            if not seg_alive[j]:
                continue
            spx, epx = d.seg_spx[b][j], d.seg_epx[b][j]
            ok = True
            for e in (spx, epx):
                ox, oy = int(e[0] * float(scale)), int(e[1] * float(scale))
                ok &= (3 <= ox < cam.width // (1 << level) - 3) and (3 <= oy < cam.height // (1 << level) - 3)
            if not ok:
                continue
            N0, dif = setup_sampling(spx, epx, d.seg_length[b][j])
            N = 1 + (N0 - 1) // (1 << level)
            P = d.seg_sf[b][j] * np.linalg.norm(d.seg_spos[b][j] - ref_pos)
            Q = d.seg_ef[b][j] * np.linalg.norm(d.seg_epos[b][j] - ref_pos)
            with np.errstate(divide="ignore", invalid="ignore"):
                inc2 = dif * float(scale) / (N - 1)
                inc3 = (Q - P) / (N - 1)
            px, X = spx * float(scale), P.copy()
            pxs, Xs = [], []
            for _ in range(N):
                pxs.append(px.copy())
                Xs.append(X.copy())
                px = px + inc2
                X = X + inc3
            seg_samples[j] = (np.array(pxs), np.array(Xs))

        def ref_cache(px, xyz):
            ui, vi, w = patch_geometry(px[:, 0], px[:, 1])
            ref = interp_block(ref_img, ui, vi, w, 0, 0)
            dx = f32(0.5) * (interp_block(ref_img, ui, vi, w, 0, 1) - interp_block(ref_img, ui, vi, w, 0, -1))
            dy = f32(0.5) * (interp_block(ref_img, ui, vi, w, 1, 0) - interp_block(ref_img, ui, vi, w, -1, 0))
            Jf = jac_xyz2uv(xyz)  # [n,2,6]
            J = (dx.astype(np.float64)[..., None] * Jf[:, None, None, 0, :] +
                 dy.astype(np.float64)[..., None] * Jf[:, None, None, 1, :]) * cJ  # [n,4,4,6]
            return ref, J

        if n_pts:
            idx = np.where(pt_fresh)[0]
            pt_ref = np.zeros((n_pts, 4, 4), f32)
            pt_J = np.zeros((n_pts, 4, 4, 6))
            if len(idx):
                px = np.stack([d.pt_px[b][idx, 0] * float(scale), d.pt_px[b][idx, 1] * float(scale)], -1)
                r, J = ref_cache(px, pt_xyz[idx])
                pt_ref[idx], pt_J[idx] = r, J
        seg_cache = {j: ref_cache(px, X) + (X,) for j, (px, X) in seg_samples.items()}

        # ---- Gauss-Newton loop (vikit NLLSSolver::optimizeGaussNewton) ----
        R_old, t_old = R.copy(), t.copy()
        for it in range(n_iter):
            H = np.zeros((6, 6))
            g = np.zeros(6)
            chi2 = 0.0
            n_meas = 0
            if n_pts:
                idx = np.where(pt_visible)[0]
                if len(idx):
                    pc = pt_xyz[idx] @ R.T + t
                    u = (cam.fx * (pc[:, 0] / pc[:, 2]) + cam.cx) * float(scale)
                    v = (cam.fy * (pc[:, 1] / pc[:, 2]) + cam.cy) * float(scale)
                    ui, vi, w = patch_geometry(u, v)
                    ok = in_frame(ui, vi, cols, rows, 2)
                    idx, ui, vi, w = idx[ok], ui[ok], vi[ok], w[ok]
                    if len(idx):
                        cur = interp_block(cur_img, ui, vi, w, 0, 0)
                        res = cur - pt_ref[idx]
                        wt = (1.0 / (1.0 + np.abs(res).astype(np.float64))).astype(f32)
                        chi2 += float(np.sum((res * res * wt).astype(np.float64)))
                        n_meas += 16 * len(idx)
                        J = pt_J[idx].reshape(-1, 6)
                        wd = wt.astype(np.float64).reshape(-1)
                        rd = res.astype(np.float64).reshape(-1)
                        H += (J * wd[:, None]).T @ J
                        g -= J.T @ (rd * wd)
            for j, (ref, J, X) in seg_cache.items():
                if not seg_alive[j]:
                    continue
                pc = X @ R.T + t
                u = (cam.fx * (pc[:, 0] / pc[:, 2]) + cam.cx) * float(scale)
                v = (cam.fy * (pc[:, 1] / pc[:, 2]) + cam.cy) * float(scale)
                ui, vi, w = patch_geometry(u, v)
                ok = in_frame(ui, vi, cols, rows, 2)
                N = len(X)
                if not ok.all():
                    seg_alive[j] = False
                    continue
                cur = interp_block(cur_img, ui, vi, w, 0, 0)
                res = cur - ref
                rho = f32(float(np.sum(np.abs(res).astype(np.float64))) / N)  # float sum order differs
                if float(rho) < 200.0:
                    wt = f32(1.0 / (1.0 + float(rho)))
                    Jf = J.reshape(-1, 6)
                    rd = res.astype(np.float64).reshape(-1)
                    H += (Jf.T @ Jf) * float(wt) / float(rho)
                    g += -(Jf.T @ rd) * float(wt)
                    chi2 += float(rho * rho * wt)
                    n_meas += 1
                else:
                    seg_alive[j] = False
            new_chi2 = float(f32(chi2) / f32(n_meas)) if n_meas else float("nan")
            try:
                x = np.linalg.solve(H, g)
            except np.linalg.LinAlgError:
                x = np.zeros(6)
            if np.isnan(x[0]):
                stop = True
            reject = (it > 0 and new_chi2 > chi2_prev) or stop
            trace.append(dict(level=level, iter=it, chi2=new_chi2, n_meas=n_meas, accepted=not reject, H=H, Jres=g, x=x))
            if reject:
                R, t = R_old, t_old
                break
            dR, dt = se3_exp(-x)
            R_old, t_old = R, t
            R, t = R @ dR, R @ dt + t
            chi2_prev = new_chi2
            if np.max(np.abs(x)) <= eps:
                break
    R_out = R @ R_ref
    t_out = R @ t_ref + t
    return dict(R=R_out, t=t_out, n_tracked=n_meas // 16, trace=trace, seg_alive=seg_alive)


# ------------------------------------------------------------------------------------------------
# pose_optimizer::optimizeGaussNewton (9-argument overload) for one frame
# ------------------------------------------------------------------------------------------------
def tukey(x):
    x = f32(x)
    b2 = f32(4.6851) * f32(4.6851)
    x2 = x * x
    if x2 <= b2:
        tmp = f32(1.0) - x2 / b2
        return float(tmp * tmp)
    return 0.0


def median_rank(v):
    v = np.sort(np.asarray(v))
    return v[len(v) // 2]


def poseopt_frame(d, b, reproj_thresh=2.0, n_iter=10):
    R, t = pose7_to_Rt(d.T_f_w[b])
    fx = d.fx
    f, pos, lvl = d.pt_f[b], d.pt_pos[b], d.pt_level[b]
    line, sp, ep, slvl = d.seg_line[b], d.seg_spos[b], d.seg_epos[b], d.seg_level[b]
    uvf = f[:, :2] / f[:, 2:3]

    def pt_err(R, t):
        pc = pos @ R.T + t
        return (uvf - pc[:, :2] / pc[:, 2:3]) / (1 << lvl)[:, None], pc

    def ln_err(R, t):
        s, e = sp @ R.T + t, ep @ R.T + t
        ds = line[:, 0] * s[:, 0] / s[:, 2] + line[:, 1] * s[:, 1] / s[:, 2] + line[:, 2]
        de = line[:, 0] * e[:, 0] / e[:, 2] + line[:, 1] * e[:, 1] / e[:, 2] + line[:, 2]
        return ds, de, s, e

    e, _ = pt_err(R, t)
    scale_pt = float(f32(1.48) * f32(median_rank(np.linalg.norm(e, axis=1).astype(f32))))
    ds, de, _, _ = ln_err(R, t)
    es, ee = ds.astype(f32), de.astype(f32)
    scale_ls = float(f32(1.48) * f32(median_rank(np.sqrt(es * es + ee * ee)))) if len(line) else 1.0
    chi2 = 0.0
    R_old, t_old = R, t
    iters = 0
    for it in range(n_iter):
        A = np.zeros((6, 6))
        bv = np.zeros(6)
        new_chi2 = 0.0
        e, pc = pt_err(R, t)
        J = jac_xyz2uv(pc) / (1 << lvl)[:, None, None]
        for i in range(len(pos)):
            w = tukey(np.linalg.norm(e[i]) / scale_pt)
            A += J[i].T @ J[i] * w
            bv -= J[i].T @ e[i] * w
            new_chi2 += e[i] @ e[i] * w
        ds, de, s, en = ln_err(R, t)
        Js, Je = jac_xyz2uv(s), jac_xyz2uv(en)
        for j in range(len(line)):
            sic = 1.0 / (1 << slvl[j])
            dsf, def_ = float(f32(ds[j])), float(f32(de[j]))
            ev = np.array([dsf, def_]) * sic
            nrm = np.linalg.norm(ev)
            k = sic * dsf / nrm
            Jl = np.stack([line[j, :2] @ (Js[j] * k), line[j, :2] @ (Je[j] * k)])
            w = tukey(nrm / scale_ls)
            A += Jl.T @ Jl * w
            bv -= Jl.T @ ev * w
            new_chi2 += ev @ ev * w
        dT = np.linalg.solve(A, bv)
        iters += 1
        if (it > 0 and new_chi2 > chi2) or np.isnan(dT[0]):
            R, t = R_old, t_old
            break
        dR, dt = se3_exp(dT)
        R_old, t_old = R, t
        R, t = dR @ R, dR @ t + dt
        chi2 = new_chi2
        if np.max(np.abs(dT)) <= 1e-10:
            break
    cov = np.linalg.inv(A * fx * fx)
    e, _ = pt_err(R, t)
    pt_out = np.linalg.norm(e, axis=1) > reproj_thresh / fx
    ds, de, _, _ = ln_err(R, t)
    el = np.stack([ds, de], -1) / (1 << slvl)[:, None]
    seg_out = np.linalg.norm(el, axis=1) > (reproj_thresh / fx) * scale_ls / scale_pt if len(line) else np.zeros(0, bool)
    return dict(R=R, t=t, cov=cov, iters=iters, pt_outlier=pt_out, seg_outlier=seg_out, scale=scale_pt * fx)
