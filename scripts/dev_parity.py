"""Developer check: GPU path vs CPU oracle at every (level, phase) stop point + full solve/CG."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import spand_public_b200 as S
import oracle_lib as O


def run(n, d, L, tol, skip=0, stops=True, verb=False):
    A = S.neglapl(n, d)
    X = S.linspace_nd(n, d)
    N = A.shape[0]
    print(f"=== n={n} d={d} L={L} tol={tol} skip={skip} N={N}")
    def mk_gpu():
        t = S.Tree(L); t.set_tol(tol); t.set_skip(skip); t.set_use_geo(True); t.set_Xcoo(X); t.set_verb(verb)
        t.partition(A); return t
    def mk_orc():
        t = O.OracleTree(L, tol=tol, skip=skip); t.set_coords(X); t.partition(A); return t
    g = mk_gpu(); o = mk_orc()
    assert (g.get_assembly_perm() == o.perm()).all()
    if stops:
        for lvl in range(min(L, 3)):
            for ph in range(4):
                g.set_stop(lvl, ph); o.set_stop(lvl, ph)
                g.assemble(A); o.partition(A); o.assemble(A)
                g.factorize(); o.factorize()
                Tg = g.get_trailing_mat(); To = o.trailing_mat()
                err = abs(Tg - To).max() if Tg.nnz + To.nnz > 0 else 0.0
                ref = abs(To).max() if To.nnz else 1.0
                sg = g.stats(); so = o.stats()
                same_rank = (sg[2] == so[2]).all()
                fg = np.sqrt((Tg.data**2).sum()); fo = np.sqrt((To.data**2).sum())
                print(f"  stop lvl {lvl} phase {ph}: max|Tg-To| = {err:.3e} (max|To| = {ref:.3e}) |T|_F {fg:.12e}/{fo:.12e} nnz {Tg.nnz}/{To.nnz} ranks equal {same_rank} fact_nnz {g.nnz()}/{o.nnz()}")
    g.set_stop(-1, -1); o.set_stop(-1, -1)
    g.assemble(A); o.partition(A); o.assemble(A)
    t0 = time.time(); g.factorize(); tg = time.time() - t0
    t0 = time.time(); o.factorize(); to = time.time() - t0
    sg = g.stats(); so = o.stats()
    ndiff = int((sg[2] != so[2]).sum())
    print(f"  factorize: gpu wall {tg:.4f}s (device {g.factorize_seconds():.4f}s, {g.kernel_launches()} launches) oracle {to:.4f}s; nnz {g.nnz()} vs {o.nnz()}; clusters with different rank: {ndiff}/{len(sg[2])}")
    lg = g.log(); lo = o.log()
    print("  dofs_left_spars gpu", lg['dofs_left_spars'].astype(int), "\n                  orc", lo['dofs_left_spars'].astype(int))
    print("  wavefronts", lg['wavefronts'].astype(int), "launches", lg['launches'].astype(int))
    print("  t_host", np.round(lg['t_host'], 4))
    for k in ('t_elim', 't_scale', 't_spars', 't_merge'):
        print(f"  {k} gpu", np.round(lg[k], 5), "orc", np.round(lo[k], 5))
    b = S.random(N, 2019)
    xg = g.solve(b); xo = o.solve(b)
    rg = np.linalg.norm(A @ xg - b) / np.linalg.norm(b); ro = np.linalg.norm(A @ xo - b) / np.linalg.norm(b)
    print(f"  one solve: res gpu {rg:.3e} orc {ro:.3e} |xg-xo|/|xo| {np.linalg.norm(xg-xo)/np.linalg.norm(xo):.3e}")
    itg, xg = g.cg(A, b, 500, 1e-12); ito, xo = o.cg(A, b, 500, 1e-12)
    print(f"  CG: gpu {itg} (t={g.t_cg:.4f}s) orc {ito}; final res gpu {np.linalg.norm(A@xg-b)/np.linalg.norm(b):.2e}")


if __name__ == "__main__":
    cases = [(5, 2, 3, 1e-2), (32, 2, 5, 1e-2), (10, 3, 4, 1e-2), (20, 2, 4, 0.0), (15, 3, 5, 1e-14)]
    if len(sys.argv) > 1:
        a = sys.argv[1:]
        run(int(a[0]), int(a[1]), int(a[2]), float(a[3]), stops=len(a) > 4 and a[4] == "stops", verb=True)
    else:
        for c in cases:
            run(*c)
        run(30, 3, 8, 1e-2, stops=False)
