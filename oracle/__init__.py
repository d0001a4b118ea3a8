"""CPU oracle for the MimIC hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch fp32 / fp64) restatement of the reference's
arithmetic for the shifted self-attention, the layer-wise alignment loss, the
LM-loss cross-entropy and the record/label-mask host logic.  Each function
cites the reference file:line it follows (paths relative to the reference
tree, mercurystraw/MimIC).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and only as the checker.  Nothing under
`mimic_b200/` imports it: the product path is the CUDA library and fails
loudly when that library is missing.

Parity pinning: the reference ships no golden vectors, so the
oracle is pinned against outputs of the reference code itself, imported by
path in the authoring container by `oracle/gen_golden.py` and committed under
`tests/golden/` (see `tests/test_oracle_golden.py`).  The LM-loss position rule
lives in `transformers==4.45.2` (not in the reference tree, not installed):
that one boundary is "parity unpinned" and stated explicitly in
`oracle/losses.py::lm_ce_loss`.
"""
