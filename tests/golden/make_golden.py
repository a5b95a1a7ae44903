"""Generate tests/golden/ref_small.npz from the UNMODIFIED reference Cython compiled into
oracle/_ref (python oracle/build_ref.py).  Run in the build container only (needs
/root/reference); the .npz is committed so the GPU box can check against it.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


def main():
    ref = oracle.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: run python oracle/build_ref.py first")
    cp, swp, clf = ref
    rs = np.random.RandomState(20261018)
    out = {}
    # ---- coding: (T,F,E) = (22,13,8), K = 7 parts of 5x5x8, thresholds on both sides of the counts
    X = (rs.rand(22, 13, 8) < 0.15).astype(np.uint8)
    X[:6] = rs.rand(6, 13, 8) < 0.02
    P = np.clip(rs.rand(7, 5, 5, 8) ** 2, .01, .99)
    lp, lip = np.log(P), np.log(1 - P)
    out.update(cp_X=X, cp_lp=lp, cp_lip=lip, cp_thr=np.int64(12))
    out["cp_out"] = cp.code_parts(X, lp, lip, 12)
    out["cp_out_fast"] = cp.code_parts_fast(X, lp, lip, 12)
    codes = np.argmax(out["cp_out"], -1)
    out["cp_codes"] = codes
    out["sp_22"] = cp.spread_patches(codes, 2, 2, 7)
    out["sp_13"] = swp.spread_waliji_patches(codes, 1, 3, 7)
    out["sp_00"] = cp.spread_patches(codes, 0, 0, 7)
    G = (rs.rand(7, 7) < 0.2).astype(np.uint8)
    Sg = out["sp_00"].copy()
    cp.spread_graph_patches(Sg, G)
    out.update(graph=G, sp_graph=Sg)
    # a non-5x5 window with E = 8 and a 3x4 part
    P2 = np.clip(rs.rand(4, 3, 4, 8) ** 2, .01, .99)
    out.update(cp2_lp=np.log(P2), cp2_lip=np.log(1 - P2))
    out["cp2_out"] = cp.code_parts(X, np.log(P2), np.log(1 - P2), 5)
    # ---- code_parts_mag_aux (code_parts.pyx:138-242): 6 magnitude bands, 3 auxiliary features
    Xm = (rs.rand(22, 6) < 0.4).astype(np.uint8)
    Xa = (rs.rand(22, 3) < 0.5).astype(np.uint8)
    tm = (rs.rand(22, 13) < 0.6).astype(np.uint8)
    Pm, Pa = np.clip(rs.rand(7, 5 + 6), .01, .99), np.clip(rs.rand(7, 5 * 3), .01, .99)
    out.update(ma_Xm=Xm, ma_Xa=Xa, ma_tm=tm, ma_lmp=np.log(Pm), ma_lmip=np.log(1 - Pm), ma_lap=np.log(Pa), ma_laip=np.log(1 - Pa))
    out["ma_out"] = cp.code_parts_mag_aux(X, Xm, Xa, lp, lip, np.log(Pm), np.log(1 - Pm), np.log(Pa), np.log(1 - Pa), tm)
    # ---- get_parts_mask (code_parts.pyx:297-348) on a single window (the reference indexes the mask with
    # absolute coordinates, so only the first window of a larger map is well defined)
    Mk = (rs.rand(5, 5) < 0.6).astype(np.uint8)
    Sx = rs.randn(5, 5)
    ps, Sp, pc, pl = cp.get_parts_mask(X[8:13, 3:8].copy(), Mk, Sx, 3)
    out.update(gp_M=Mk, gp_S=Sx, gp_patches=ps, gp_Sp=Sp, gp_counts=pc, gp_locs=pl)
    # ---- detection: F (40,6,5), L = 9
    F = (rs.rand(40, 6, 5) < 0.2).astype(np.uint8)
    LF = rs.randn(9, 6, 5).astype(np.float32)
    out.update(dt_F=F, dt_LF=LF, dt_out=clf.detect(F, LF))
    Ff = rs.rand(40, 6, 5).astype(np.float32)
    out.update(df_F=Ff, df_out=clf.detect_float(Ff, LF))
    LF4 = rs.randn(3, 9, 6, 5).astype(np.float32)
    cs = rs.randn(3).astype(np.float32)
    mx, am = clf.detect_float_max(Ff, LF4, cs)
    out.update(dm_LF=LF4, dm_cs=cs, dm_max=mx, dm_arg=am)
    F2 = rs.rand(40, 11).astype(np.float32)
    LF2 = rs.randn(9, 11).astype(np.float32)
    out.update(d2_F=F2, d2_LF=LF2, d2_out=clf.detect_float2d(F2, LF2), d2_spec=clf.detect_spectral(F2, LF2))
    F4 = (rs.rand(3, 25, 6, 5) < 0.2).astype(np.uint8)
    LF64 = rs.randn(9, 6, 5)
    lengths = np.array([25, 10, 0], np.int64)
    out.update(de_F=F4, de_LF=LF64, de_len=lengths, de_max=clf.get_max_detect_examples(F4, LF64, lengths))
    # ---- clustering of detection peaks (cluster_times.pyx:133-219), compiled reference
    ct = oracle.load_ref_cluster_times()
    if ct is None:
        raise SystemExit("oracle/_ref/cluster_times is not built: run python oracle/build_ref.py first")
    gaps = [rs.choice([1, 2, 3, 5, 8, 13, 30, 60], size=n) for n in (0, 1, 2, 7, 19, 40, 40, 64)]
    for i, g in enumerate(gaps):
        t = np.cumsum(g).astype(np.uint16)
        C0, C1 = float([3, 3, 5, 5, 10, 3, 33, 8][i]), float([10, 10, 20, 33, 50, 20, 100, 15][i])
        out["ct_times_%d" % i], out["ct_C_%d" % i] = t, np.array([C0, C1], np.float32)
        out["ct_part_%d" % i] = ct.cluster_times(t, C0, C1)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_small.npz"), **out)
    print("wrote ref_small.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
