"""CPU oracle for the DiffPure reverse-SDE hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain PyTorch CPU fp32, the arithmetic of the reference path
(NVlabs/DiffPure: runners/diffpure_sde.py, runners/diffpure_guided.py, runners/diffpure_ddpm.py and
the three score-model UNets they call). Each function cites the reference file:line it follows.

Rules (enforced by tests/test_layout.py):
  * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it;
  * nothing under diffpure_b200/ imports it -- the product path is the CUDA library and fails loudly
    when the library is missing.

Pinning: the reference ships no golden vectors or tests for this path, so the
oracle is pinned against the reference's own Python modules imported from /root/reference in the build
container (oracle/check_against_reference.py, oracle/make_golden.py); the resulting vectors are
committed under tests/golden/ and re-checked on every run (tests/test_oracle_golden.py).
The Euler-Maruyama stepping itself lives in the third-party `torchsde` package (unpinned in
diffpure.Dockerfile:62, absent offline): its published fixed-step Ito-Euler rule is restated in
oracle/sde.py and anchored on the reference's call site runners/diffpure_sde.py:228-239 and on the
in-repo EulerMaruyamaPredictor (score_sde/sampling.py:177-187) -- for that one piece parity is
"unpinned against torchsde itself" and says so in DESIGN.md.
"""
