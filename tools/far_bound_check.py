#!/usr/bin/env python
"""Adversarial check of the far-field bound of 580-raytracer_b200/csrc/fargrid.cuh against the float arithmetic of the
reference's triangle test (Raytracer.cpp:392, 937-942).

Claim: with u = 2^-24, q = P - v1, e = v2 - v1, m = e x N, a plane hit P on the OUTER side of the edge v1v2 passes the first
area test (da >= 0 in float) only if
        |q.m| <= 6.0001 u G(q, N) + 9.3 u |q| |e|,      G(q, N) = |qy qz Nx| + |qz qx Ny| + |qx qy Nz| <= 0.57741 |q|^2
so that for w = q / |q| acceptance needs |q| >= L(w) = (|w.m| - 9.3 u |e|) / (6.0001 u g(w, N)).

The script places P at kappa * L(w) along in-wedge directions w (slightly tilted out of the plane, as real rays are) for random
triangles of the benchmark scenes' size range, evaluates da with numpy float32 in the reference's operation order, and counts
accepted outer points.  kappa < 1 must give ZERO accepts (the bound is necessary); kappa > 1 gives accepts (it is not vacuous).
Run: python tools/far_bound_check.py [samples_per_kappa]      (CPU only, ~1 min for the default 4e6 per kappa)
"""
import sys
import numpy as np
f32 = np.float32; u = 2.0**-24
rng = np.random.default_rng(7)
def cross(a, b): return np.stack([a[:,1]*b[:,2]-a[:,2]*b[:,1], a[:,2]*b[:,0]-a[:,0]*b[:,2], a[:,0]*b[:,1]-a[:,1]*b[:,0]], 1)
def dot(a, b): return (a[:,0]*b[:,0] + a[:,1]*b[:,1]) + a[:,2]*b[:,2]
def run(n, kappa, tilt=1e-4):
    c = rng.uniform(-100, 100, (n, 3))
    size = 10.0**rng.uniform(-2.5, 1, (n, 1))
    v0 = (c + size*rng.normal(size=(n,3))).astype(f32); v1 = (c + size*rng.normal(size=(n,3))).astype(f32); v2 = (c + size*rng.normal(size=(n,3))).astype(f32)
    e1 = v1 - v0; e2 = v2 - v0
    Nn = cross(e1, e2); ln = np.sqrt(dot(Nn, Nn)); N = (Nn/ln[:,None]).astype(f32)
    total = f32(0.5)*dot(cross(e1, e2), N)
    # double precision geometry
    V0, V1, V2, Nd = v0.astype(np.float64), v1.astype(np.float64), v2.astype(np.float64), N.astype(np.float64)
    E1, E2 = V1 - V0, V2 - V0
    e = V2 - V1; m = np.cross(e, Nd)
    # in-plane wedge direction: a E1 + b E2 with a,b>0
    a = rng.uniform(0, 1, (n,1)); w = a*E1/np.linalg.norm(E1,axis=1,keepdims=True) + (1-a)*E2/np.linalg.norm(E2,axis=1,keepdims=True)
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    w = w + tilt*rng.normal(size=(n,1))*Nd
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    g = np.abs(w[:,1]*w[:,2]*Nd[:,0]) + np.abs(w[:,2]*w[:,0]*Nd[:,1]) + np.abs(w[:,0]*w[:,1]*Nd[:,2])
    wm = np.abs((w*m).sum(1)); el = np.linalg.norm(e, axis=1)
    Lmin = (wm - 9.3*u*el)/(6.0001*u*g)          # direction-aware
    Lmin0 = (wm - 9.3*u*el)/(6.0001*u*0.57741)   # direction-independent g bound
    out = {}
    for name, LL in (('aware', Lmin), ('indep', Lmin0)):
        L = kappa*LL
        P = (V1 + L[:,None]*w).astype(f32)
        da = dot(cross(v1 - P, v2 - P), N)
        with np.errstate(all='ignore'):
            neg = (f32(0.5)*da)/total < 0
        # true side
        S = -((P.astype(np.float64) - V1)*m).sum(1)
        outer = (S*total.astype(np.float64)) < 0
        ok = np.isfinite(L) & (L > 0) & (L < 1e30) & outer
        out[name] = (int(ok.sum()), int((ok & ~neg).sum()))
    return out
reps = max(1, (int(float(sys.argv[1])) if len(sys.argv) > 1 else 4000000) // 400000)
bad = 0
for kappa in (0.5, 0.9, 0.99, 1.5, 5, 30):
    tot = {'aware':[0,0], 'indep':[0,0]}
    for rep in range(reps):
        r = run(400000, kappa)
        for k in r: tot[k][0] += r[k][0]; tot[k][1] += r[k][1]
    print('kappa', kappa, tot)
    if kappa < 1:
        bad += tot['aware'][1] + tot['indep'][1]
print('accepts below the bound:', bad, '(must be 0)')
sys.exit(1 if bad else 0)
