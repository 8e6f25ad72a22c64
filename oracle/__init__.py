"""CPU oracle for the segger hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-torch (CPU) + scipy restatement of the reference
algorithms on the hot path named by BASELINE.json (GATv2 hetero message
passing, tx<->cell scoring, kNN graph construction).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import it -- as the checker / baseline, never as
the product.  Nothing under ``segger_b200/`` imports this package.

PARITY STATUS
-------------
* kNN sub-path (``oracle.neighbors``): PINNED -- it issues the very same
  ``scipy.spatial.KDTree(points, leafsize=100).query(k, distance_upper_bound,
  workers=-1)`` call the reference makes
  (/root/reference/src/segger/data/utils/neighbors.py:139-150) and restates
  ``knn_to_edge_index`` (:54-92) 1:1.
* GATv2 / HeteroConv / scatter_max sub-paths: **PARITY UNPINNED**.  The
  arithmetic lives in torch_geometric 2.7.0 and torch_scatter 2.1.2
  (pixi.lock:3408,3470) which are absent from /root/reference and not
  installable here; the reference ships no tests or golden vectors.  The
  restatement follows SURVEY.md Appendix A and is self-checked (fp64
  gradcheck, dense-attention equivalence, permutation equivariance).
* The segger-side torch code (sinusoidal embedding, positional embedder,
  ISTEncoder.forward plumbing, predict_step) is restated 1:1 from
  models/ist_encoder.py and models/lightning_model.py.
"""
