"""CPU oracle for the NSC hot path -- TEST INFRASTRUCTURE ONLY.

This package is a behavioural restatement (numpy / scipy / torch-CPU) of the reference's
hot path.  It is the checker, never the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import it;
  * nothing under ``nsc_b200/`` imports it, and the product fails loudly when the CUDA
    library is missing instead of falling back to this code.

PARITY STATUS: **parity unpinned**.  The reference ships no tests, golden vectors, checkpoints
or data, and its arithmetic lives in TensorFlow / audiolazy / spectrum, none of which is
installed here, so the oracle cannot be pinned against the reference's
own outputs.  What *is* pinned (tests/test_oracle_pins.py):
  * spectrum's published poly2lsf / lsf2poly doc-string example (the MATLAB known answer),
  * the reference's literal constants (256 LSF bins, init_alpha, filter taps, window sums),
  * closed-form identities (analysis->synthesis round trip, Parseval for the rFFT, mel-matrix
    partition of unity, SAME-padding table of SURVEY.md section 3.2).
Every behaviour taken from third-party library knowledge is marked [LIB] next to the code.
"""
