"""Developer script: step-by-step parity of the CUDA path against the oracle (run under gpurun)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qgb200
import qg_oracle as o


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def mk(M, P, dt=3600.0, kick=1e-6):
    mo = o.standard_model(M, P, dt=dt, initial_kick=kick)
    mg = qgb200.BaroclinicModel(mo.H_1, mo.H_2, mo.beta, mo.Lx, mo.Ly, mo.dt, mo.T, mo.U, mo.M, mo.P, mo.dx,
                                mo.visc, mo.r, mo.R_d, mo.initial_kick)
    return mo, mg


def check(M, P, nsteps=10, backend="spectral"):
    mo, mg = mk(M, P)
    zeta, psi = o.initialise_model(mo, seed=1)
    f = np.zeros_like(zeta)
    fac = o.make_factors(mo, backend)
    zg, pg, fg = zeta.copy(order="F"), psi.copy(order="F"), f.copy(order="F")
    with qgb200.Session(mg) as s:
        s.upload(zg, pg, fg)
        # roundtrip
        z2, p2, f2 = s.new_state_array(), s.new_state_array(), s.new_state_array()
        s.download(z2, p2, f2)
        print(f"[{M}x{P}] roundtrip", rel(z2, zeta), rel(p2, psi))
        # single evolve_zeta
        zo, po, fo = zeta.copy(order="F"), psi.copy(order="F"), f.copy(order="F")
        o.evolve_zeta(mo, zo, po, 1, fo)
        s.evolve_zeta(1)
        s.download(z2, p2, f2)
        print(f"[{M}x{P}] evolve_zeta(1): q {rel(z2[:,:,:,0], zo[:,:,:,0]):.2e} f {rel(f2[:,:,:,0], fo[:,:,:,0]):.2e} hist {rel(z2, zo):.2e}")
        o.evolve_psi(mo, zo, po, *fac)
        s.evolve_psi()
        s.download(z2, p2, f2)
        for l in range(2):
            print(f"[{M}x{P}] evolve_psi layer{l}: {rel(p2[:,:,l,0], po[:,:,l,0]):.2e}")
        print(f"[{M}x{P}] psi hist {rel(p2, po):.2e}")
        # continue to nsteps
        o.run_steps(mo, zo, po, fo, fac, 2, nsteps - 1)
        s.step(2, nsteps - 1)
        s.download(z2, p2, f2)
        print(f"[{M}x{P}] after {nsteps} steps: psi {rel(p2[:,:,:,0], po[:,:,:,0]):.2e} q {rel(z2[:,:,:,0], zo[:,:,:,0]):.2e} "
              f"f {rel(f2, fo):.2e} hist psi {rel(p2, po):.2e} q {rel(z2, zo):.2e}")
        Eo, Zo = o.diagnostics(mo, zo, po)
        Eg, Zg = s.diagnostics()
        print(f"[{M}x{P}] E rel {abs(Eg-Eo)/Eo:.2e} Z rel {abs(Zg-Zo)/Zo:.2e}")


if __name__ == "__main__":
    sizes = [(8, 8), (16, 8), (64, 64), (24, 40), (128, 128), (256, 512), (1024, 1024)]
    for M, P in sizes:
        try:
            check(M, P, backend="direct" if M * P <= 128 * 128 else "spectral")
        except Exception as e:
            import traceback; traceback.print_exc()
    # throughput quick look
    for M in (1024, 4096):
        mo, mg = mk(M, M, dt=300.0)
        z, p = qgb200.initialise_model(mg, seed=1)
        f = np.zeros_like(z)
        with qgb200.Session(mg) as s:
            s.upload(z, p, f)
            s.step(1, 5); s.sync()
            t = time.time(); s.step(6, 50); s.sync(); dt = time.time() - t
            print(f"[{M}] {50*M*M/dt:.3e} cell-steps/s  {dt/50*1e3:.3f} ms/step")
            s.set_profiling(True); s.step(56, 20); s.sync()
            for k, (ms, n) in s.kernel_times().items():
                if n: print(f"   {k:16s} {ms/n*1e3:9.1f} us x{n}")
            print("   E,Z", s.diagnostics())
