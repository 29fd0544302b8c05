"""CPU oracle for the conv-GAT hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker or as the timed
CPU baseline -- never as the thing shipped.  The product path
(``extended-gan_b200/``) never imports this package and raises when its CUDA
library is missing.

Contents
--------
``spec.py``        plain-PyTorch (fp32/fp64, CPU) restatement of the reference
                   arithmetic: adjacency normalisation, the per-pixel graph
                   attention of ``convolutional_gat/baseline_model.py``, the 1-D
                   layer, the builder's spec of the missing ``GATMultiHead3D``
                   (SURVEY.md appendix A.2), SmaAt-UNet (public architecture) and
                   the DCGAN nets.  Every function cites the reference file:line
                   it follows.
``ref_loader.py``  imports the UNMODIFIED in-tree reference modules from
                   ``/root/reference`` (container only; the GPU box has no
                   reference) behind an ``ipdb`` stub and a no-op ``.cuda`` shim.
``make_golden.py`` generates ``tests/golden/*.pt`` from the live reference.

Parity status
-------------
PINNED (against the live in-tree reference, via tests/golden and
tests/test_oracle_vs_reference.py): adjacency normalisation, GraphAttentionLayer2D
/ GATMultiHead2D / BaselineModel2D, GraphAttentionLayer / GATMultiHead /
BaselineModel, DCGAN Generator / FrameDiscriminator / TemporalDiscriminator.
PARITY UNPINNED: ``GATMultiHead3D``, ``GATMultistream.Model`` and ``SmaAt_UNet``
-- their source is absent from the reference tree (un-vendored ``GAT3D``
sub-module with no recorded URL or revision).  The spec here degenerates to the
pinned in-tree layers (linear mapping + pixel soft-max == GraphAttentionLayer2D),
and SmaAt-UNet is pinned only by its parameter count (4,032,548,
``compare_models/results/results.json:18``).
"""
