"""High-precision known answers for the monotonic rational-quadratic spline (best-effort pin of the spline path).

zuko is absent from /root/reference and from the image, so no reference-held golden vector exists for this boundary
("parity unpinned"). What CAN be pinned independently of every floating-point implementation in this repo is
the closed form itself: Durkan et al., "Neural Spline Flows" (NeurIPS 2019), eqs. (4)-(8), with the knot
parameterisation zuko's MonotonicRQSTransform documents (zuko >= 1.0, `zuko/transforms.py`; soft clip
a -> a / (1 + |f a / log slope|) with f = 2 for widths / heights and f = 1 for derivatives, softmax -> cumulative sum ->
[-B, B], derivatives exp(d) with 1 at both ends), evaluated here in 50-digit mpmath arithmetic from hand-set
unconstrained parameters. The answers are stored as decimal strings and as the nearest doubles.

    python tests/golden/make_rqs_known_answers.py      ->  tests/golden/rqs_known_answers.json
"""
import json
import os

import mpmath as mp

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))


def spline(w, h, d, bound, slope):
    """knots (X, Y) and knot derivatives D as mpf lists of length K + 1"""
    K = len(w)
    if slope is not None:
        L = mp.log(mp.mpf(slope))
        w = [a / (1 + abs(2 * a / L)) for a in w]
        h = [a / (1 + abs(2 * a / L)) for a in h]
        d = [a / (1 + abs(a / L)) for a in d]
    def knots(a):
        e = [mp.e ** v for v in a]
        s = mp.fsum(e)
        c, out = mp.mpf(0), [-mp.mpf(bound)]
        for v in e:
            c += v / s
            out.append(mp.mpf(bound) * (2 * c - 1))
        return out
    X, Y = knots(w), knots(h)
    D = [mp.mpf(1)] + [mp.e ** v for v in d] + [mp.mpf(1)]
    assert len(D) == K + 1
    return X, Y, D


def forward(X, Y, D, x):
    K = len(X) - 1
    if not (X[0] < x <= X[K]):            # searchsorted(right=False) - 1 outside [0, K): identity
        return x, mp.mpf(0), (-1 if x <= X[0] else K)
    k = max(i for i in range(K) if X[i] < x)
    dx, dy = X[k + 1] - X[k], Y[k + 1] - Y[k]
    s = dy / dx
    z = (x - X[k]) / dx
    den = s + (D[k] + D[k + 1] - 2 * s) * z * (1 - z)
    y = Y[k] + dy * (s * z * z + D[k] * z * (1 - z)) / den
    jac = s * s * (D[k + 1] * z * z + 2 * s * z * (1 - z) + D[k] * (1 - z) ** 2) / den ** 2
    return y, mp.log(jac), k


def inverse(X, Y, D, y):
    K = len(X) - 1
    if not (Y[0] < y <= Y[K]):
        return y, (-1 if y <= Y[0] else K)
    k = max(i for i in range(K) if Y[i] < y)
    dx, dy = X[k + 1] - X[k], Y[k + 1] - Y[k]
    s = dy / dx
    t = y - Y[k]
    a = dy * (s - D[k]) + t * (D[k] + D[k + 1] - 2 * s)
    b = dy * D[k] - t * (D[k] + D[k + 1] - 2 * s)
    c = -s * t
    z = 2 * c / (-b - mp.sqrt(b * b - 4 * a * c))
    return X[k] + z * dx, k


CASES = [
    dict(name="K5_bound1", bound=1.0, slope=1e-3,
         w=[0.3, -1.2, 0.7, 2.0, -0.4], h=[-0.8, 0.1, 1.5, -2.2, 0.6], d=[0.9, -1.1, 0.2, 1.7]),
    dict(name="K8_bound5_large_parameters", bound=5.0, slope=1e-3,
         w=[4.0, -7.5, 0.0, 12.0, -3.3, 0.25, 6.0, -0.5], h=[-9.0, 2.2, 2.2, -0.1, 15.0, -4.0, 0.7, 1.0],
         d=[3.0, -8.0, 0.5, 20.0, -0.6, 1.1, -2.4]),
    dict(name="K3_no_clip", bound=2.0, slope=None, w=[0.5, -0.25, 1.0], h=[1.25, 0.0, -0.75], d=[0.4, -0.9]),
    dict(name="K1", bound=1.1, slope=1e-3, w=[0.7], h=[-0.3], d=[]),
]


def main():
    out = []
    for case in CASES:
        w, h, d = ([mp.mpf(repr(v)) for v in case[k]] for k in ("w", "h", "d"))
        X, Y, D = spline(w, h, d, case["bound"], case["slope"])
        K = len(w)
        B = mp.mpf(repr(case["bound"]))
        xs = [-B * mp.mpf("1.3"), -B * mp.mpf("0.999"), B * mp.mpf("0.9999"), B * mp.mpf("1.2"), mp.mpf(0)]
        for k in range(K):                      # two points inside every bin
            xs += [X[k] + (X[k + 1] - X[k]) * mp.mpf("0.37"), X[k] + (X[k + 1] - X[k]) * mp.mpf("0.91")]
        pts = []
        for x in xs:
            x = mp.mpf(repr(float(x)))          # the input the kernels will actually see: a double
            y, ladj, k = forward(X, Y, D, x)
            xi, ki = inverse(X, Y, D, x)        # the same number read as an ordinate
            _, ladj_at_xi, _ = forward(X, Y, D, xi)
            pts.append(dict(x=float(x), y=mp.nstr(y, 30), ladj=mp.nstr(ladj, 30), bin=k,
                            x_inv=mp.nstr(xi, 30), bin_inv=ki, ladj_at_x_inv=mp.nstr(ladj_at_xi, 30)))
        out.append(dict(name=case["name"], bound=case["bound"], slope=case["slope"], widths=case["w"], heights=case["h"],
                        derivatives=case["d"], knots_x=[mp.nstr(v, 30) for v in X], knots_y=[mp.nstr(v, 30) for v in Y],
                        knot_derivatives=[mp.nstr(v, 30) for v in D], points=pts))
    with open(os.path.join(HERE, "rqs_known_answers.json"), "w") as f:
        json.dump(dict(source="Durkan et al. 2019 eqs. (4)-(8) + zuko >= 1.0 MonotonicRQSTransform parameterisation, mpmath 50 digits",
                       cases=out), f, indent=1)
    print("written:", sum(len(c["points"]) for c in out), "points in", len(out), "cases")


if __name__ == "__main__":
    main()
