"""The identity behind the round-1 Lagrange-basis commitments (csrc/prover.cu, pk_build_lagrange / lagrange_commit), on scalars.

With the test SRS a KZG commitment is p(tau) * G, so "the Lagrange-basis MSM equals the coefficient-basis commitment" is an identity
between field elements: the oracle's own first round (oracle/marlin_oracle.py prove(), following ark-marlin 0.3.0 ahp/prover.rs
prover_first_round) builds w and z_A as coefficient vectors; their values at tau must equal the sums the product forms from the basis
scalars  L_k(tau)  and  (L_k(tau) - [k in X] l_j^X(tau)) / v_X(tau)  and the small integers it feeds the kernels.  CPU only."""
import random

import numpy as np
import pytest

from oracle import marlin_oracle as mo

P = mo.P


def basis_scalars(dh, dx, tau):
    """what pk_build_lagrange computes on the device: (L_k(tau))_k and ((L_k(tau) - [k in X] l_j(tau)) / v_X(tau))_k"""
    h, x = dh.size, dx.size
    vh, vx = dh.vanishing(tau), dx.vanishing(tau)
    u = [pow(dh.gen, k, P) * pow((tau - pow(dh.gen, k, P)) % P, -1, P) % P for k in range(h)]
    c_l = vh * pow(h, -1, P) % P
    c_w = c_l * pow(vx, -1, P) % P
    lag = [c_l * uk % P for uk in u]
    lagw = [c_w * uk % P for uk in u]
    ratio = h // x
    for j in range(x):
        lagw[j * ratio] = (c_w - pow(x, -1, P)) * u[j * ratio] % P
    return lag, lagw, vh, vx


@pytest.mark.parametrize("log_h,log_x", [(6, 3), (7, 1), (5, 4)])
def test_w_and_z_commitments_in_the_lagrange_basis(log_h, log_x):
    rng = random.Random(1000 * log_h + log_x)
    dh, dx = mo.Domain(1 << log_h), mo.Domain(1 << log_x)
    h, x = dh.size, dx.size
    ratio = h // x
    tau = rng.randrange(2, P)
    inst = [1] + [rng.randrange(2) for _ in range(x - 1)]          # formatted public input (bits, leading one)
    wit = [rng.randrange(2) for _ in range(h - x - 3)]              # a few padding zeros at the end, as pad_r1cs leaves
    r_w, r_a = rng.randrange(P), rng.randrange(P)

    # ---- the oracle's first round, w polynomial (oracle/marlin_oracle.py:508-524) ----
    x_poly = dx.ifft(mo.vec_to_m(inst))
    x_evals = dh.fft(x_poly)
    w_ext = wit + [0] * (h - x - len(wit))
    kk = np.arange(h)
    src = kk - kk // ratio - 1
    w_vals = mo.vec_to_m(w_ext + [0])[np.where(kk % ratio == 0, len(w_ext), src)]
    w_evals = mo.vsub(w_vals, x_evals)
    w_evals[kk % ratio == 0] = 0
    w_full = mo._add_vanishing_multiple(dh.ifft(w_evals), mo.to_m(r_w), h)
    w_poly, rem = mo.divide_by_vanishing(w_full, x)
    assert len(rem) == 0
    lag, lagw, vh, vx = basis_scalars(dh, dx, tau)

    # ---- the product's form: the full assignment in H order (po_assignment_h_i32) against the w basis ----
    a = [0] * h
    for k in range(h):
        if k % ratio == 0:
            a[k] = inst[k // ratio]
        else:
            wi = k - k // ratio - 1
            a[k] = wit[wi] if wi < len(wit) else 0
    got = (sum(ak * bk for ak, bk in zip(a, lagw)) + r_w * vh * pow(vx, -1, P)) % P
    assert got == mo.poly_eval(w_poly, tau)

    # ---- z_A: small signed row sums against the plain Lagrange basis ----
    za = [rng.randrange(-2, 3) for _ in range(h)]
    z_a_poly = mo._add_vanishing_multiple(dh.ifft(mo.vec_to_m([v % P for v in za])), mo.to_m(r_a), h)
    got = (sum(v * lk for v, lk in zip(za, lag)) + r_a * vh) % P
    assert got == mo.poly_eval(z_a_poly, tau)
