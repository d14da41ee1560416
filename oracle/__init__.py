"""CPU oracle for the PointNet++ hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product package
(``khairil_tum-facade_semantic_segmentation_b200``) never does.
"""
