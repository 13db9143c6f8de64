"""Host build of the device scene generator (csrc/scenegen.h, compiled with g++): sha256 sub-seeds, SeedSequence ->
PCG64, Generator.integers / uniform draws against NumPy itself, and whole scenes against the host generator
(carlabev_env_b200/scenes.py, which is bit-identical to the reference's post-reset state).  Everything drawn or
derived without smoothing must be EXACT; the smoothed routes agree to 1e-9 px (see scenegen.h)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAXA, MAXR = 3, 8


class Actor(C.Structure):
    _fields_ = [("kind", C.c_int), ("n", C.c_int), ("beh", C.c_int), ("tidx0", C.c_int),
                ("state0", C.c_double * 4), ("cruise_px", C.c_double), ("cruise_mps", C.c_double),
                ("beh_p", C.c_double * 4), ("cx", C.c_double * MAXR), ("cy", C.c_double * MAXR),
                ("cyaw", C.c_double * MAXR), ("raw_x", C.c_double * MAXR), ("raw_y", C.c_double * MAXR)]


class Scene(C.Structure):
    _fields_ = [("ego_state0", C.c_double * 4), ("ego_target_speed", C.c_double), ("len_ego_route", C.c_double),
                ("ego_tidx0", C.c_int), ("num_vehicles", C.c_int), ("n_actors", C.c_int), ("attempts", C.c_int),
                ("ego_cx", C.c_double * 6), ("ego_cy", C.c_double * 6), ("ego_cyaw", C.c_double * 6),
                ("rew_rx", C.c_int32 * 6), ("rew_ry", C.c_int32 * 6), ("rew_cum", C.c_double * 6),
                ("actors", Actor * MAXA)]


def _ang_close(a, b, tol=1e-9):
    d = (np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64) + np.pi) % (2 * np.pi) - np.pi
    return bool(np.all(np.abs(d) < tol))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("sgh") / "libsgh.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "scenegen_host.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.sgh_derive_seed.restype = C.c_int64
    lib.sgh_derive_seed.argtypes = [C.c_int64, C.c_char_p]
    lib.sgh_draws.argtypes = [C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    lib.sgh_generate.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(Scene)]
    assert lib.sgh_scene_bytes() == C.sizeof(Scene)
    return lib


def test_derived_seeds_and_generator_draws_match_numpy(lib):
    from carlabev_env_b200.scenes import derive_seed

    for base in (0, 1, 7, 4095, 123456789, 2**31 - 2, 2**40 + 17):
        for part in ("route", "traffic", "scenario"):
            assert lib.sgh_derive_seed(base, part.encode()) == derive_seed(base, part)
    for seed in (0, 1, 5, 2**31 - 2, 1234567890123, 2**63 + 11):
        for lo, hi in ((900, 1000), (-1, 2), (0, 7), (5, 6)):
            n = 64
            uni = np.zeros(n)
            ints = np.zeros(n, dtype=np.int64)
            lib.sgh_draws(seed, n, uni.ctypes.data, ints.ctypes.data, lo, hi)
            g = np.random.default_rng(seed)
            for i in range(n):
                assert int(g.integers(lo, hi)) == ints[i], (seed, lo, hi, i)
                assert float(g.uniform(-2.0, 5.0)) == uni[i], (seed, i)


@pytest.mark.parametrize("kind,levels", [("lead_brake", (1, 2, 3)), ("jaywalk", (1, 2, 3, 4))])
def test_generated_scenes_match_the_host_generator(lib, kind, levels):
    from carlabev_env_b200.engine import savgol_operators
    from carlabev_env_b200.scenes import build_scripted_scene
    from golden_util import load_map

    cls = load_map()
    sg = np.ascontiguousarray(savgol_operators())
    kid = {"lead_brake": 1, "jaywalk": 2}[kind]
    retried = 0
    for seed in list(range(60)) + [4095, 99991, 2**31 - 5]:
        for level in levels:
            ref = build_scripted_scene(kind, seed, level=level, cls_map=cls)
            s = Scene()
            ok = lib.sgh_generate(kid, level, seed, sg.ctypes.data, cls.ctypes.data, cls.shape[1], cls.shape[0], 182,
                                  C.byref(s))
            assert ok == 1
            retried += s.attempts > 1
            na = len(ref["act_kind"])
            assert s.n_actors == na and s.num_vehicles == int(ref["num_vehicles"])
            # exact: everything drawn or derived without the smoothing stage
            assert list(s.rew_rx) == ref["rew_rx"].tolist() and list(s.rew_ry) == ref["rew_ry"].tolist()
            assert s.ego_target_speed == float(ref["ego_target_speed"]) and s.len_ego_route == float(ref["len_ego_route"])
            assert s.ego_tidx0 == int(ref["ego_tidx0"]) and s.ego_state0[3] == ref["ego_state0"][3]
            ro, wo = ref["act_route_off"], ref["act_raw_off"]
            for a in range(na):
                A = s.actors[a]
                assert A.kind == int(ref["act_kind"][a]) and A.beh == int(ref["act_beh"][a])
                assert A.n == ro[a + 1] - ro[a] == wo[a + 1] - wo[a]
                assert list(A.beh_p) == ref["act_beh_p"][a].tolist()
                assert A.cruise_mps == ref["act_cruise_mps"][a] and A.cruise_px == ref["act_cruise_px"][a]
                assert list(A.raw_x)[:A.n] == ref["act_raw_x"][wo[a]:wo[a + 1]].tolist()
                assert list(A.raw_y)[:A.n] == ref["act_raw_y"][wo[a]:wo[a + 1]].tolist()
                assert A.tidx0 == int(ref["act_tidx0"][a]) and A.state0[3] == ref["act_state0"][a][3]
                # smoothed: SciPy's LAPACK edge fit vs the linear operator
                for mine, key in ((A.cx, "act_cx"), (A.cy, "act_cy")):
                    assert np.allclose(list(mine)[:A.n], ref[key][ro[a]:ro[a + 1]], rtol=0, atol=1e-9), (seed, level, key)
                # headings modulo 2 pi: a route that runs due west has yaw = +pi or -pi depending on the SIGN of the
                # 1e-13 px smoothing noise in its constant coordinate
                assert _ang_close(list(A.cyaw)[:A.n], ref["act_cyaw"][ro[a]:ro[a + 1]]), (seed, level, "act_cyaw")
                assert np.allclose(list(A.state0)[:2], ref["act_state0"][a][:2], rtol=0, atol=1e-9)
                assert _ang_close([A.state0[2]], [ref["act_state0"][a][2]])
                # the spawn jitter itself is an exact integer on top of the smoothed start point
                assert round(A.state0[0] - A.cx[0]) == round(ref["act_state0"][a][0] - ref["act_cx"][ro[a]])
            for mine, key in ((s.ego_cx, "ego_cx"), (s.ego_cy, "ego_cy")):
                assert np.allclose(list(mine), ref[key], rtol=0, atol=1e-9), (seed, level, key)
            assert _ang_close(list(s.ego_cyaw), ref["ego_cyaw"])
            assert np.allclose(list(s.ego_state0)[:2], ref["ego_state0"][:2], rtol=0, atol=1e-9)
            assert _ang_close([s.ego_state0[2]], [ref["ego_state0"][2]])
    assert retried >= 0
