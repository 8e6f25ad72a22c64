"""CPU oracle for the segger hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-torch (CPU) + numpy + scipy restatement of the reference algorithms on the hot path named by
BASELINE.json (GATv2 hetero message passing, tx<->cell scoring, kNN graph construction) and of the rows SURVEY 8f marks
"next" (losses, prediction graph, tiling, writer post-processing).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it -- as the checker / baseline, never as the
product.  Nothing under ``segger_b200/`` imports this package.

PARITY STATUS (round 2)
-----------------------
PINNED to the reference's own code.  ``oracle/reference_import.py`` executes the reference's hot-path files UNMODIFIED
from /root/reference (``models/ist_encoder.py``, ``models/triplet_loss.py``, ``models/lightning_model.py``,
``data/utils/neighbors.py``, ``data/partition/sampler.py``, ``io/fields.py``) under a synthetic package with
``sys.modules`` stand-ins for the third-party names they import.  ``tests/golden/make_reference_golden.py`` ran them on
seeded inputs and committed the outputs (``tests/golden/ref_*.pt|npz``); ``tests/test_reference_pin.py`` checks the
oracle against those fixtures everywhere, and re-runs the reference live (incl. fixture freshness and signature checks)
wherever /root/reference exists.  Covered this way: sinusoidal / positional embedder, ISTEncoder / SkipGAT forward and
gradients (generic H=3 and the bench's H=2 shapes, train-mode dropout through injected masks), FastTripletSelector /
TripletLoss / MetricLoss, LitISTEncoder.get_losses / weight schedule / predict_step, kdtree_neighbors /
knn_to_edge_index / setup_transcripts_graph / setup_prediction_graph, the bin-packing samplers.

Still UNPINNED (third-party arithmetic that is absent from /root/reference and not installable here; each is a
restatement of the published algorithm and says so in its module header):
* the inside of PyG 2.7.0 ``GATv2Conv`` / ``HeteroConv`` / ``HeteroDictLinear`` and ``torch_scatter.scatter_max``
  (``oracle/pyg_ref.py``, wrapped in the PyG API by ``oracle/pyg_stub.py`` for the reference's files to call): SURVEY
  Appendix A, self-checked by fp64 gradcheck, a dense masked-attention formulation, permutation equivariance;
* cuSpatial's point-in-polygon join (``oracle/geometry_ref.py``), skimage's Yen / Li thresholds
  (``oracle/writer_ref.py``), PyG ``HeteroData.subgraph`` / collate (``oracle/tiles_ref.py``).
* kNN: PINNED since round 1 -- the oracle issues the very ``scipy.spatial.KDTree(...).query`` call the reference makes
  (data/utils/neighbors.py:139-150).
"""
