"""CPU oracle for the WeatherConverter hot path (TEST INFRASTRUCTURE ONLY).

A plain-PyTorch fp32 restatement of the reference's DDPM reverse-sampling loop with
semantic-gradient guidance.  Every function cites the reference file:line it follows.

Rules (enforced by tests/test_host_logic.py::test_no_oracle_in_product):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
    ``--impl reference`` legs may import anything from this package;
  * nothing under ``weatherconverter_b200/`` imports it; the product path raises when the
    CUDA library is missing instead of falling back to this code.

Parity status: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference's own modules, generated in the build
container by ``tests/golden/make_golden.py`` (which imports ``/root/reference``) and
committed under ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks them.
"""
