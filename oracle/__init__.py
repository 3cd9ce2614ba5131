"""CPU oracle for the NSC hot path -- TEST INFRASTRUCTURE ONLY.

This package is a behavioural restatement (numpy / scipy / torch-CPU) of the reference's
hot path.  It is the checker, never the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import it;
  * nothing under ``nsc_b200/`` imports it, and the product fails loudly when the CUDA
    library is missing instead of falling back to this code.

PARITY STATUS: pinned against the REFERENCE'S OWN SOURCE FILES executed in the build container, with the third-party
libraries it imports replaced by stand-ins -- not against TensorFlow itself.  The reference ships no tests, golden vectors,
checkpoints or data, and TensorFlow / audiolazy / spectrum are not installable here (no network), so its code cannot run as
shipped.  tests/golden/make_ref_golden.py imports the unmodified utilities.py, lpc_utilities.py, nn_core_operator.py,
loss_terms_and_measures.py and neural_speech_coding_module.py from /root/reference and runs their function bodies -- window
constructions, frame loops, sub-frame weighting, the codec graph of all four topologies (layer order, shapes, activations,
residuals, sub-pixel permutes), quantiser, loss terms -- on behavioural stand-ins written independently of this package
(tests/golden/tf_shim.py: torch-backed conv / softmax / top_k / stft / mel matrix; scipy lfilter / solve_toeplitz / numpy.roots for
audiolazy and spectrum).  The outputs are committed (tests/golden/reference_run*.npz); tests/test_reference_run_pins.py checks
this oracle against them and re-runs the generator where /root/reference exists; the -m gpu tests check the CUDA path against the
same vectors.  What stays an assumption is the SEMANTICS of the third-party calls (marked [LIB]): TF 'SAME' padding and variable
creation order, top_k's tie rule, audiolazy's zero-state filters and autocorrelation LPC, spectrum's LSF conventions -- each now
with a second, independent implementation behind it.  Also pinned (tests/test_oracle_pins.py, test_oracle_second_source.py):
  * spectrum's published poly2lsf / lsf2poly doc-string example (the MATLAB known answer),
  * the reference's literal constants (256 LSF bins, init_alpha, filter taps, window sums),
  * closed-form identities (analysis->synthesis round trip, Parseval for the rFFT, mel-matrix
    partition of unity, SAME-padding table of SURVEY.md section 3.2),
  * second sources: torchaudio's HTK mel matrix and lfilter, numpy's Toeplitz solve, torch's SAME conv.
Every behaviour taken from third-party library knowledge is marked [LIB] next to the code.
"""
