"""GPU parity of the writer post-processing against oracle/writer_ref.py: de-duplication bit-exact,
thresholds within 1e-6 relative (floating point; scikit-image restated, see the oracle's header)."""
import numpy as np
import pytest
import torch

from oracle import writer_ref
from segger_b200 import writer

pytestmark = pytest.mark.gpu


def _predictions(n=40000, n_genes=60, n_cells=500, seed=0, dup_frac=0.05):
    rng = np.random.default_rng(seed)
    row = rng.permutation(n).astype(np.int64)
    gene = rng.integers(0, n_genes, n).astype(np.int32)
    gene[gene == 7] = 8                                     # gene 7 absent
    seg = rng.integers(0, n_cells, n).astype(np.int64)
    seg[rng.random(n) < 0.2] = -1
    # bimodal similarities per gene (what the thresholds separate), float32
    sim = np.where(rng.random(n) < 0.6, rng.normal(0.75, 0.08, n), rng.normal(0.3, 0.1, n)).clip(-1, 1).astype(np.float32)
    sim[seg < 0] = 0.0
    gene_of_row = gene.copy()
    # duplicates: some transcripts predicted twice (tile borders), the copies keep the gene
    k = int(n * dup_frac)
    pick = rng.choice(n, k, replace=False)
    row2, gene2 = row[pick], gene_of_row[pick]
    seg2 = rng.integers(-1, n_cells, k).astype(np.int64)
    sim2 = rng.uniform(0, 1, k).astype(np.float32)
    sim2[:50] = sim[pick[:50]]                              # exact similarity ties -> lowest cell encoding wins
    sim2[seg2 < 0] = 0.0
    parts = [(row[:n // 2], seg[:n // 2], sim[:n // 2], gene[:n // 2]), (row[n // 2:], seg[n // 2:], sim[n // 2:], gene[n // 2:]),
             (row2, seg2, sim2, gene2)]
    return parts, n_genes


def test_dedupe_bit_exact_vs_oracle():
    parts, _ = _predictions()
    cat = [np.concatenate([p[i] for p in parts]) for i in range(4)]
    want = writer_ref.dedupe_ref(*cat)
    got = writer.dedupe_predictions(*[torch.from_numpy(c).cuda() for c in cat])
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    assert np.array_equal(got[0].cpu().numpy(), np.arange(40000))
    # 64-bit row indices (two radix passes over the key) and an empty input
    big = cat[0] + (1 << 35)
    got = writer.dedupe_predictions(torch.from_numpy(big).cuda(), *[torch.from_numpy(c).cuda() for c in cat[1:]])
    assert np.array_equal(got[0].cpu().numpy(), np.arange(40000) + (1 << 35)) and np.array_equal(got[1].cpu().numpy(), want[1])
    e = writer.dedupe_predictions(*[torch.from_numpy(c[:0]).cuda() for c in cat])
    assert all(t.numel() == 0 for t in e)


def test_gene_thresholds_vs_oracle_incl_absent_constant_and_unconverged():
    parts, n_genes = _predictions(seed=1)
    row, seg, sim, gene = writer_ref.dedupe_ref(*[np.concatenate([p[i] for p in parts]) for i in range(4)])
    sim = sim.copy()
    sim[(gene == 3) & (seg >= 0)] = 0.5                      # constant gene: threshold = that value
    want = writer_ref.gene_thresholds_ref(gene, seg, sim)
    thr, conv, counts = writer.gene_thresholds(torch.from_numpy(gene).cuda(), torch.from_numpy(seg).cuda(),
                                               torch.from_numpy(sim).cuda(), n_genes)
    thr, conv, counts = thr.cpu().numpy(), conv.cpu().numpy(), counts.cpu().numpy()
    assert counts[7] == 0 and np.isnan(thr[7]) and not conv[7]
    assert thr[3] == pytest.approx(0.5, abs=1e-7)
    assert set(np.nonzero(counts)[0].tolist()) == set(want)
    for g, (t, c) in want.items():
        assert conv[g] == c, g
        assert thr[g] == pytest.approx(t, rel=1e-6, abs=1e-7), g
        assert counts[g] == int(((gene == g) & (seg >= 0)).sum())
    # iteration cap: with max_iter = 2 most genes fail to converge and are back-filled with the median of the rest
    thr2, conv2, _ = writer.gene_thresholds(torch.from_numpy(gene).cuda(), torch.from_numpy(seg).cuda(),
                                            torch.from_numpy(sim).cuda(), n_genes, max_iter=2)
    ref2 = {}
    failed = []
    for g in want:
        arr = sim[(gene == g) & (seg >= 0)]
        try:
            ref2[g] = min(writer_ref.threshold_yen(arr) if arr.max() > arr.min() else float(arr[0]), writer_ref.threshold_li(arr, 2))
        except writer_ref.NotConverged:
            failed.append(g)
    assert failed and ref2
    glob = float(np.quantile(list(ref2.values()), 0.5))
    thr2, conv2 = thr2.cpu().numpy(), conv2.cpu().numpy()
    for g in failed:
        assert not conv2[g] and thr2[g] == pytest.approx(glob, rel=1e-6)
    for g, t in ref2.items():
        assert conv2[g] and thr2[g] == pytest.approx(t, rel=1e-6, abs=1e-7)


def test_assign_transcripts_to_cells_columns_and_parquet(tmp_path):
    parts, n_genes = _predictions(n=12000, seed=2)
    ids = [f"cell-{i}" for i in range(500)]
    want = writer_ref.assign_transcripts_to_cells_ref(parts, ids)
    got = writer.assign_transcripts_to_cells([[torch.from_numpy(a) for a in p] for p in parts], ids, n_genes=n_genes)
    assert np.array_equal(got["row_index"], want["row_index"])
    assert list(got["segger_cell_id"]) == list(want["segger_cell_id"])
    assert np.array_equal(got["segger_similarity"], want["segger_similarity"])
    has = want["has_threshold"]
    assert np.allclose(got["similarity_threshold"][has], want["similarity_threshold"][has], rtol=1e-6, atol=1e-7)
    assert np.isnan(got["similarity_threshold"][~has]).all()
    assert np.array_equal(got["converged"][has], want["converged"][has])
    import pyarrow.parquet as pq
    path = tmp_path / "segger_segmentation.parquet"
    writer.write_segmentation(got, path)
    t = pq.read_table(path)
    assert t.column_names == ["row_index", "segger_cell_id", "segger_similarity", "similarity_threshold", "converged"]
    assert t.num_rows == 12000 and t.column("segger_cell_id").null_count == int((np.array([c is None for c in got["segger_cell_id"]])).sum())
