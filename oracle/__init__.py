"""CPU oracle for the sdflabel render/refine hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``sdflabel_b200``) never imports anything from here and fails loudly when its
CUDA library is missing.

What it is: an independent torch-CPU restatement (fp32 or fp64, autograd used as
the differentiation engine) of the algorithm the reference implements in

    sdfrenderer/grid.py, sdfrenderer/deepsdf/networks/deep_sdf_decoder_scale.py,
    sdfrenderer/deepsdf/workspace.py, sdfrenderer/renderer/{projection,primitives,
    rasterer,utils_rasterer}.py, pipelines/optimizer.py, utils/refinement.py:108-125

Every function cites the reference file:line it follows.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this
path, so the oracle is pinned against outputs of the
reference itself, run unmodified in the build container: ``oracle/make_golden.py``
imports the reference from /root/reference, drives it and this restatement with
the same seeded inputs and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
re-checks the restatement against those files on every run (no reference needed).
"""
