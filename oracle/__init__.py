"""oracle/ — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch fp32 / fp64) restatement of the reference's policy-training hot path
(thobotics/geometry_rl: HEPi / EMPN / transformer actor, DeepSets critic, GAE, TRPL projection,
TRPL loss).  Every function cites the reference file:line it follows (paths relative to the
reference checkout, `/root/reference` in the build container).

Who may import this package: `tests/`, `__graft_entry__.smoke()` (through `smoke_check.py`, as the checker) and
`bench.py`'s `cpu_baseline` / `--impl reference` legs.  Nothing under `geometry_rl_b200/` imports
it; the product path raises if its CUDA library is missing instead of falling back to this code.

Pinning status
--------------
* Model bodies (HEPi, EMPN/Ponita, DeepSets, transformer), Gaussian head, mean / W2 projection,
  KL helpers, trust-region loss and metrics: **pinned** against outputs of the *unmodified*
  reference modules imported from `/root/reference` through `oracle/ref_shims`
  (`oracle/make_golden.py` -> `tests/golden/*.pt`, checked by `tests/test_oracle_golden.py`).
  The shims restate torch-geometric 2.5.2 / torch_scatter / torch_cluster semantics from memory
  (SURVEY.md "[3P-memory]"), so "pinned" means pinned to the reference's own source lines, with the
  third-party primitives restated.
* KL covariance projection (ITPAL `cpp_projection`, unpinned upstream, not installable offline),
  torchrl 0.3.1 `GAE` and torch_cluster kNN tie-breaking: **parity unpinned** — restated from the
  published algorithm; validated by KKT residuals, fp64 gradcheck and a naive reverse loop.
"""
