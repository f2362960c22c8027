#!/usr/bin/env python
"""How far does a build with FMA contraction drift from the strict (bit-exact) build?  Measurement aid only.

    python tools/fma_variant_check.py run OUT.npz [--poly-order K]      # with GORILLA_B200_LIB unset (strict) or set (variant)
    python tools/fma_variant_check.py compare STRICT.npz VARIANT.npz

`run` pushes seeded particles of the analytic-tokamak workload for one time step through the loaded library and stores
the end state; `compare` reports the share of particles that end in the same tetrahedron with the same number of pushes and
the largest relative deviation of position and parallel velocity.  The strict build is the product (the
visited-tetra sequence is only reproducible without contraction); the variant exists to put a number on what that costs.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("run")
    r.add_argument("out")
    r.add_argument("--poly-order", type=int, default=4)
    r.add_argument("--n", type=int, default=100000)
    r.add_argument("--t-step", type=float, default=1e-4)
    c = sub.add_parser("compare")
    c.add_argument("strict")
    c.add_argument("variant")
    a = ap.parse_args()
    if a.cmd == "run":
        import workloads
        from gorilla_b200 import Gorilla, build_mesh
        grid, st = workloads.analytic_tokamak(40, 40, 40)
        st.poly_order = a.poly_order
        g = Gorilla(build_mesh(grid, st), st)
        x, vpar, vperp = workloads.particles_cyl(a.n, 7)
        binit, ind, ifc = workloads.fresh_state(a.n)
        npush = np.zeros(a.n, np.int64)
        g.orbit_timestep_gorilla(x, vpar, vperp, a.t_step, binit, ind, ifc, n_pushes=npush)
        e, pphi, _ = g.invariants(x, vpar, vperp, ind)
        np.savez(a.out, x=x, vpar=vpar, vperp=vperp, ind=ind, npush=npush, energy=e, pphi=pphi)
        print(f"{a.out}: {int(npush.sum())} pushes, kernel {g.counters().kernel_ms:.2f} ms")
        return
    s, v = np.load(a.strict), np.load(a.variant)
    ok = (s["ind"] > 0) & (v["ind"] > 0)
    rel = lambda p, q: float(np.max(np.abs(p - q) / np.maximum(np.abs(p), 1e-300)))  # noqa: E731
    print(json.dumps({
        "particles": int(ok.size), "both_inside": int(ok.sum()),
        "same_final_tetra": float(np.mean(s["ind"] == v["ind"])),
        "same_push_count": float(np.mean(s["npush"] == v["npush"])),
        "pushes_strict": int(s["npush"].sum()), "pushes_variant": int(v["npush"].sum()),
        "max_rel_dev_R": rel(s["x"][ok, 0], v["x"][ok, 0]), "max_rel_dev_vpar": rel(s["vpar"][ok], v["vpar"][ok]),
        "max_rel_dev_energy": rel(s["energy"][ok], v["energy"][ok]),
    }))


if __name__ == "__main__":
    main()
