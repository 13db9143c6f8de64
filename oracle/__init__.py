"""CPU oracle for the CarlaBEV batched-stepping hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain Python/NumPy, the algorithm of the reference's
per-step path.  Every function cites the reference
file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product
package `carlabev_env_b200` never does.

Parity status
-------------
* Dynamics, scripted actors, collision / target ordering, CaRL + shaping rewards,
  episode statistics and the mask / frame-stack wrappers are PINNED: the
  restatement is checked bit-for-bit (float64 equality on this CPU) against golden
  trajectories produced by the UNMODIFIED reference modules stepped in the build
  container (oracle/gen_golden.py; fixtures in tests/golden/), and fuzzed in lock-step
  with the reference under random configurations, reset options and actions
  (oracle/fuzz_steps.py, oracle/fuzz_scenes.py; results in DESIGN.md section 2).
* The reference renders with pygame 2.6.1 and wraps with gymnasium 1.2.2, neither
  installed here.  The golden run therefore sits on oracle/shims (a restatement of
  those third-party libraries).  Raster values (transform.rotate sampling, rect
  rounding, blit clipping) and wrapper values are therefore "PARITY UNPINNED"
  beyond the reference's own contracts (anchor pixel == hero colour for several
  yaws, crop alignment, shapes), which tests/test_oracle_contracts.py replays.
  cv2.resize(INTER_AREA) IS installed and is used directly as the pin for the
  resize stage.
"""
