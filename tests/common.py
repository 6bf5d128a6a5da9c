"""Shared test configurations (seeded, small enough for the CPU checkers to finish in seconds)."""
import numpy as np

EL = lambda E, nu=0.3: (0, E, nu, 0.0, 0.0, 0.0)
PL = lambda E, nu, Ka, Sy: (1, E, nu, Ka, Sy, 0.0)
DM = lambda E, nu, Xt: (2, E, nu, 0.0, 0.0, Xt)

# micro-structures / materials used across tests
CASES = {
    # config 2 of BASELINE.json, shrunk: elastic sphere, contrast 10
    "elastic_sphere": dict(type=1, geo_params=(0.2, 0.0, 0.0, 0.0), materials=[EL(1e7), EL(1e8), EL(1e7)]),
    # config 3: damage matrix + elastic sphere (test/benchmark-sc2019.cpp:84-86)
    "damage_sphere": dict(type=1, geo_params=(0.2, 0.0, 0.0, 0.0), materials=[DM(1e7, 0.3, 1e5), EL(3e7), EL(3e7)]),
    # config 4: elastic matrix + plastic layer / fibre (test/test3d_4.cpp:65-67)
    "plastic_layer": dict(type=2, geo_params=(0.5, 0.0, 0.0, 0.0),
                          materials=[EL(1e6), PL(1e3, 0.3, 5e4, 1e3), EL(1e6)]),
    "plastic_fibre": dict(type=4, geo_params=(0.2, 0.0, 0.0, 0.0),
                          materials=[EL(1e6), PL(1e3, 0.3, 5e4, 1e3), EL(1e6)]),
    # the same micro-structures with the reference's golden plastic material (test/benchmark-plastic.cpp:78), the one
    # bench.py's plastic40 workload uses: E = 3e7, Sy = 1e5 yields at strains of a few 1e-3, so the J2 return mapping,
    # its forward-difference tangent and the state-variable update are really exercised (E = Sy = 1e3 above never yields)
    "plastic_layer_yield": dict(type=2, geo_params=(0.5, 0.0, 0.0, 0.0),
                                materials=[EL(3e7, 0.25), PL(3e7, 0.25, 1e7, 1e5), EL(3e7, 0.25)]),
    "plastic_fibre_yield": dict(type=4, geo_params=(0.2, 0.0, 0.0, 0.0),
                                materials=[EL(3e7, 0.25), PL(3e7, 0.25, 1e7, 1e5), EL(3e7, 0.25)]),
    "homog_damage": dict(type=0, materials=[DM(1e7, 0.3, 1e5), EL(1e7), EL(1e7)]),
    "mic3d_8": dict(type=10, materials=[EL(1e7), EL(1e8), DM(5e6, 0.3, 1e5)]),
}


def relerr(a, b, floor=1e-300):
    """max-norm error relative to the max-norm of the reference."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor))


def random_u(nndim, seed, scale=1e-3):
    rng = np.random.default_rng(seed)
    return rng.uniform(-scale, scale, nndim)


def random_vars(nelem, mat_types_by_elem, seed):
    """Plausible internal variables in the reference layout [e][gp][7]."""
    rng = np.random.default_rng(seed)
    v = np.zeros((nelem, 8, 7))
    for e in range(nelem):
        t = mat_types_by_elem[e]
        if t == 1:  # plastic: eps_p (small), alpha >= 0
            v[e, :, :6] = rng.uniform(-1e-4, 1e-4, (8, 6))
            v[e, :, 6] = rng.uniform(0, 1e-4, 8)
        elif t == 2:  # damage: r, D
            v[e, :, 0] = rng.uniform(20.0, 60.0, 8)
            v[e, :, 1] = rng.uniform(0.0, 0.5, 8)
    return v.reshape(-1)
