"""CPU oracle for the compress/decompress hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.

Pinning status
--------------
* Analysis / synthesis transforms (all convolution arithmetic, activations,
  residual adds, image I/O casts): PINNED.  ``oracle/cae_oracle.py`` is a
  restatement of ``/root/reference/src/models/tasks/_autoencoders.py`` that is
  checked bit-for-bit against the reference's own classes imported here
  (``oracle/ref_loader.py``) by ``oracle/make_golden.py`` and by
  ``tests/test_oracle_vs_reference.py`` (skipped where ``/root/reference`` is
  absent); the outputs are committed under ``tests/golden/``.
* EntropyBottleneck, GDN, ``pmf_to_quantized_cdf`` and the rANS coder live in
  the un-vendored third-party dependency ``compressai>=1.2.4``
  (``/root/reference/requirements.txt:26``), which is absent from this
  container and cannot be installed (no network).  They are restated from the
  published algorithm.  **Parity unpinned** for these
  pieces: they are anchored only by analytic known-answer tests, round trips
  and the reference's own call sites (``_autoencoders.py:476-502,549-572``).
"""
