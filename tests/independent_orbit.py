"""An orbit integrator that shares NOTHING with the pusher restatements (oracle/, gorilla_b200/csrc) except the per-tetrahedron
record: TEST INFRASTRUCTURE, the "second opinion" of tests/test_independent_pin.py.

Inside a tetrahedron the guiding-centre equations of GORILLA are the linear system dz/dtau = b + A z, z = (x - x1, v_par)
(written by the surveyor from SRC/pusher_tetra_poly.f90:125-178,1503-1530 -- a different reader of
the Fortran than the author of the oracle).  The pusher solves it with a truncated Taylor series and a polynomial root
solver for the exit time; here it is integrated numerically with scipy's DOP853 (rtol 1e-13) and the exit through a face is
located by event detection on the dense output; because that interpolant is only good to ~1e-9, the event time is then
polished with Brent's method on the EXACT flow of the linear system (matrix exponential of the augmented matrix, Pade +
scaling and squaring -- again no Taylor series in tau and no polynomial root solver).  Hand-over to the neighbour (SRC/pusher_tetra_func_mod.f90:6-93, kind 1) and
the time accounting t = tau * dt_dtau_const are restated in numpy.  No series, no root solver, no fall-back ladder.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp
from scipy.linalg import expm
from scipy.optimize import brentq

CLIGHT = 2.9979e10   # constants_mod.f90
# offsets (doubles) into type tetrahedron_physics, SRC/tetra_physics_mod.f90:9-83
X1, DIST_REF, ANORM, CURLA, BMOD1, PHI1, DTDTAU = 0, 3, 9, 21, 24, 30, 40
GBXCURLA, GPHIXCURLA, SPALP, SPBET, GBXH1, GPHIXH1, GB, GPHI, CURLH, ALP, BET = 41, 42, 47, 48, 50, 53, 59, 62, 89, 107, 116
# type tetrahedron_grid, SRC/tetra_grid_mod.f90:6-15
NB_TETR, NB_FACE, PER_PHI, PER_THETA = 4, 8, 12, 16


def _ode(rec, z0, perpinv, cm_over_e, sign_rhs):
    """b (4), A (4x4) of the push that starts at z0 ('init' and 'coeffs')."""
    gB, gPhi = rec[GB:GB + 3], rec[GPHI:GPHI + 3]
    bmod0 = rec[BMOD1] + gB @ z0[:3]
    phi0 = rec[PHI1] + gPhi @ z0[:3]
    vperp2 = -2.0 * perpinv * bmod0
    k1 = vperp2 + z0[3] ** 2 + 2.0 * perpinv * rec[BMOD1]
    k3 = rec[PHI1] - phi0
    curlh = rec[CURLH:CURLH + 3]
    b = np.empty(4)
    b[:3] = (curlh * k1 + perpinv * rec[GBXH1:GBXH1 + 3]) * cm_over_e - CLIGHT * (2.0 * k3 * curlh + rec[GPHIXH1:GPHIXH1 + 3])
    b[3] = perpinv * rec[GBXCURLA] - CLIGHT / cm_over_e * rec[GPHIXCURLA]
    A = np.zeros((4, 4))
    alp = rec[ALP:ALP + 9].reshape(3, 3).T     # Fortran column-major alpmat(i,j)
    bet = rec[BET:BET + 9].reshape(3, 3).T
    A[:3, :3] = perpinv * cm_over_e * alp - CLIGHT * bet
    A[3, 3] = perpinv * cm_over_e * rec[SPALP] - CLIGHT * rec[SPBET]
    A[:3, 3] = rec[CURLA:CURLA + 3]
    return b * sign_rhs, A * sign_rhs


def independent_orbit(mesh, x, vpar, vperp, ind_tetr, t_step, max_crossings=400, rtol=1e-13):
    """Follow one particle (start INSIDE tetrahedron ind_tetr, 1-based) for the physical time t_step.
    Returns dict(seq = [(ind_tetr, iface) after every push], x, vpar, vperp, ind_tetr, margin, complete) where margin is the
    smallest relative distance of an exit point from an edge of its exit face (small = a marginal crossing, where a pusher
    of finite order may legitimately pick the neighbouring face) and complete says whether t_step was consumed."""
    tp, tg, sc = mesh.tetra_physics, mesh.tetra_grid, mesh.scalars
    cm, coord = sc["cm_over_e"], sc["coord_system"]
    per_phi, per_theta = 2.0 * np.pi / sc["n_field_periods"], 2.0 * np.pi
    sign_rhs = sc["sign_sqg"] * (-1 if t_step < 0 else 1)
    x = np.array(x, float)
    rec = tp[ind_tetr - 1]
    z = np.append(x - rec[X1:X1 + 3], vpar)
    perpinv = -0.5 * vperp ** 2 / (rec[BMOD1] + rec[GB:GB + 3] @ z[:3])
    t_remain, iface_in, seq, margin = float(t_step), 0, [], np.inf
    for _ in range(max_crossings):
        rec = tp[ind_tetr - 1]
        z = np.append(x - rec[X1:X1 + 3], z[3])
        an = rec[ANORM:ANORM + 12].reshape(4, 3)     # anorm(:,f) = row f
        off = np.array([rec[DIST_REF], 0.0, 0.0, 0.0])
        b, A = _ode(rec, z, perpinv, cm, sign_rhs)
        dt_dtau = rec[DTDTAU] * sign_rhs
        tau_stop = t_remain / dt_dtau                # > 0: the time step ends inside this cell if no face comes first

        def rhs(_, y):
            return b + A @ y

        def face_event(f):
            def g(_, y):
                return an[f] @ y[:3] + off[f]
            g.terminal, g.direction = True, -1.0     # leaving: the (inward) normal distance falls through 0
            return g
        d0 = an @ z[:3] + off
        scale = np.abs(d0).max()
        events = [face_event(f) for f in range(4)]
        y0 = z.copy()
        if iface_in:
            # start exactly ON the entry face (the hand-over leaves a rounding-size distance of either sign)
            n = an[iface_in - 1]
            y0[:3] -= n * (d0[iface_in - 1] / (n @ n))
        M = np.zeros((5, 5))
        M[:4, :4], M[:4, 4] = A, b
        y0a = np.append(y0, 1.0)

        def flow(tau):                               # exact solution of dz/dtau = b + A z
            return (expm(M * tau) @ y0a)[:4]
        # step bound: ~30 steps per cell transit, so that a face distance cannot dip below zero and come back between two
        # steps unseen (the distances are nearly parabolic in tau)
        tau_cell = scale / max(np.abs(an @ (b + A @ y0)[:3]).max(), 1e-300)
        sol = solve_ivp(rhs, (0.0, tau_stop), y0, method="DOP853", rtol=rtol, atol=1e-300, events=events,
                        first_step=min(tau_stop, 1e-3 * tau_cell), max_step=tau_cell / 30.0)
        hit = [f for f in range(4) if len(sol.t_events[f])]
        if not hit:                                  # time step consumed inside the cell
            zf = flow(tau_stop)
            x = zf[:3] + rec[X1:X1 + 3]
            seq.append((ind_tetr, 0))
            vp = np.sqrt(2.0 * abs(perpinv) * (rec[BMOD1] + rec[GB:GB + 3] @ zf[:3]))
            return dict(seq=seq, x=x, vpar=zf[3], vperp=vp, ind_tetr=ind_tetr, margin=margin, complete=True)
        f = min(hit, key=lambda k: sol.t_events[k][0])
        tau = sol.t_events[f][0]
        gf = lambda t: an[f] @ flow(t)[:3] + off[f]  # noqa: E731
        lo, hi, w, tries = tau, tau, 1e-7 * tau, 0
        while gf(lo) <= 0.0 and lo > 0.0 and tries < 40:
            lo, w, tries = max(0.0, lo - w), 4.0 * w, tries + 1
        w = 1e-7 * tau
        while gf(hi) > 0.0 and tries < 80:
            hi, w, tries = hi + w, 4.0 * w, tries + 1
        if not (lo < hi) or gf(lo) <= 0.0 or gf(hi) > 0.0 or (iface_in == f + 1 and tau < 1e-6 * tau_stop):
            # no clean sign change around the detected event, or the orbit leaves at once through the face it came in by
            # (it turns ON the face: the pusher's "prolonged trajectory" business): a marginal crossing, nothing to compare
            return dict(seq=seq, x=x, vpar=z[3], vperp=np.nan, ind_tetr=ind_tetr, margin=0.0, complete=False)
        tau = brentq(gf, lo, hi, xtol=1e-300, rtol=8.9e-16)
        zf = flow(tau)
        others = np.delete(an @ zf[:3] + off, f)
        margin = min(margin, others.min() / scale)   # how far inside the other three faces the exit point lies
        # belt and braces: on the exact flow no face may have been crossed before tau
        for tq in np.linspace(0.0, tau, 34)[1:-1]:
            dq = an @ flow(tq)[:3] + off
            if iface_in:
                dq[iface_in - 1] = max(dq[iface_in - 1], 0.0) if tq < 0.1 * tau else dq[iface_in - 1]
            if dq.min() < -1e-12 * scale:
                margin = 0.0
        t_remain -= tau * dt_dtau
        x = zf[:3] + rec[X1:X1 + 3]
        z = zf
        g = tg[ind_tetr - 1]
        nxt, iface_in = int(g[NB_TETR + f]), int(g[NB_FACE + f])
        pphi, pth = int(g[PER_PHI + f]), int(g[PER_THETA + f])
        iphi = 1 if coord == 1 else 2
        x[iphi] -= pphi * per_phi
        if coord == 2:
            x[1] -= pth * per_theta
        seq.append((nxt, iface_in))
        vp = np.sqrt(2.0 * abs(perpinv) * (rec[BMOD1] + rec[GB:GB + 3] @ zf[:3]))
        if nxt < 1:                                  # left the domain
            return dict(seq=seq, x=x, vpar=zf[3], vperp=vp, ind_tetr=-1, margin=margin, complete=False)
        ind_tetr = nxt
    return dict(seq=seq, x=x, vpar=z[3], vperp=vp, ind_tetr=ind_tetr, margin=margin, complete=False)
