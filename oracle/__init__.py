"""CPU oracle for the graph-convolution hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the timed CPU baseline.
The product path (``gnn_tableextraction_b200``) never imports it and fails
loudly when the CUDA library is missing.

What it restates: ``/root/reference/src/components/graphs/models.py:15-170``
(``GcnSAGELayer``, ``GcnSAGE``, ``WeightedMeanSAGELayer``, ``MeanSAGE``) and the
train-step envelope ``/root/reference/src/models/model_train.py:320-332``.

Parity status.  The arithmetic the reference delegates to DGL
(``update_all(u_mul_e, sum|mean)``, ``in_degrees``, ``dgl.batch``) lives in a
third-party dependency that is neither vendored under /root/reference nor pinned
by it (requirements.txt / setup.py list no dgl version; API usage points at DGL
0.6-0.9), and DGL cannot be installed here (no network).  The reference ships no
tests, golden vectors or fixtures.  Therefore:

  * the layer composition (concat order, norm, LayerNorm, activation, init) IS
    pinned: ``tests/golden/make_golden.py`` imports the reference's own
    ``models.py`` unmodified, runs it over ``oracle/dgl_shim.py`` and commits
    the resulting vectors under ``tests/golden/``; the oracle is checked
    against them bit-for-bit-level tight (fp32, same op order);
  * the DGL primitives themselves are restated from DGL's published semantics
    (``oracle/dgl_shim.py`` header lists each one) and are NOT pinned by any
    reference-side vector: **parity unpinned at the DGL kernel level**.  Each
    assumed semantic has its own unit test in ``tests/test_oracle.py`` so a
    later cross-check against a real DGL install can falsify it cheaply.
"""
