"""Runs the reference's OWN Python files for the hot path, unmodified, in this container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  `/root/reference` exists only in the build
container: nothing on the GPU box may call this module -- it is used by
`tests/golden/make_reference_golden.py` (fixture generator) and by the `-m "not gpu"` pin tests,
which skip when the reference tree is absent.

How: segger's package `__init__` files import cupy / rmm / cuspatial / polars / lightning (absent,
SURVEY.md section 0.3), but the files ON the hot path are plain torch (+ scipy) on top of a few
third-party classes.  Each file is executed from where it lies with `importlib` under a synthetic
package (`_segger_ref`), with `sys.modules` stubs for exactly the third-party names it imports:

  models/ist_encoder.py      torch_geometric.nn           -> oracle.pyg_stub (PyG API over oracle.pyg_ref math)
  models/triplet_loss.py     torch_geometric.data         -> placeholder classes (imported, never used)
  models/lightning_model.py  lightning, torch_scatter, polars, ..io.fields (the real file), ..data.data_module (stub)
  data/partition/sampler.py  torch_geometric.loader, .dataset (placeholders): the bin-packing functions are pure Python
  data/utils/neighbors.py    geopandas, polars, cupy, cugraph, cuml, cudf (placeholders); ...geometry.points_in_polygons
                             -> oracle.geometry_ref; scipy.spatial.KDTree is the real one

The stubs are removed from `sys.modules` again before `load()` returns.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path
from types import SimpleNamespace

REF_ROOT = Path(os.environ.get("SEGGER_REFERENCE", "/root/reference")) / "src" / "segger"
_PKG = "_segger_ref"
_cache = None


def available() -> bool:
    return (REF_ROOT / "models" / "ist_encoder.py").exists()


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _package(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = []          # a package: relative imports resolve through sys.modules
    return m


def _exec(name: str, path: Path) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load() -> SimpleNamespace:
    """-> namespace(ist_encoder, triplet_loss, lightning_model, neighbors, fields): the reference modules."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT} (it exists only in the build container)")
    from . import geometry_ref, pyg_stub
    from .ist_encoder_ref import scatter_max_ref

    class _Placeholder:
        def __init__(self, *a, **k):
            raise NotImplementedError("placeholder for a third-party class the hot path never instantiates")

    def _points_in_polygons(points, polygons, predicate="contains", batches=1):
        raise NotImplementedError("cuSpatial join: see oracle.geometry_ref.points_in_polygons_ref")

    stubs = {
        "torch_geometric": _package("torch_geometric"),
        "torch_geometric.nn": _module("torch_geometric.nn", GATv2Conv=pyg_stub.GATv2Conv, Linear=pyg_stub.Linear,
                                      HeteroDictLinear=pyg_stub.HeteroDictLinear, HeteroConv=pyg_stub.HeteroConv),
        "torch_geometric.data": _module("torch_geometric.data", Data=_Placeholder, Batch=_Placeholder,
                                        HeteroData=_Placeholder),
        "torch_geometric.loader": _module("torch_geometric.loader", DataLoader=_Placeholder, DynamicBatchSampler=_Placeholder),
        "torch_scatter": _module("torch_scatter", scatter_max=lambda src, index, dim_size=None: scatter_max_ref(
            src, index, int(dim_size))),
        "lightning": _module("lightning", LightningModule=pyg_stub.LightningModule),
        "polars": _module("polars", DataFrame=_Placeholder, Expr=_Placeholder, Series=_Placeholder),
        "geopandas": _module("geopandas", GeoDataFrame=_Placeholder, GeoSeries=_Placeholder),
        "cupy": _module("cupy"), "cugraph": _module("cugraph"), "cuml": _module("cuml"), "cudf": _module("cudf"),
        # synthetic package tree for the reference's relative imports
        _PKG: _package(_PKG),
        f"{_PKG}.models": _package(f"{_PKG}.models"),
        f"{_PKG}.io": _package(f"{_PKG}.io"),
        f"{_PKG}.data": _package(f"{_PKG}.data"),
        f"{_PKG}.data.utils": _package(f"{_PKG}.data.utils"),
        f"{_PKG}.geometry": _module(f"{_PKG}.geometry", points_in_polygons=_points_in_polygons),
        f"{_PKG}.data.partition": _package(f"{_PKG}.data.partition"),
        f"{_PKG}.data.partition.dataset": _module(f"{_PKG}.data.partition.dataset", PartitionDataset=_Placeholder),
        f"{_PKG}.data.data_module": _module(f"{_PKG}.data.data_module", ISTDataModule=type("ISTDataModule", (), {})),
    }
    saved = {k: sys.modules.get(k) for k in stubs}
    added = []
    try:
        sys.modules.update(stubs)
        fields = _exec(f"{_PKG}.io.fields", REF_ROOT / "io" / "fields.py")
        added.append(f"{_PKG}.io.fields")
        for n in ("StandardBoundaryFields", "TrainingBoundaryFields", "StandardTranscriptFields",
                  "TrainingTranscriptFields"):
            setattr(sys.modules[f"{_PKG}.io"], n, getattr(fields, n))
        mods = {}
        for key, rel in (("ist_encoder", "models/ist_encoder.py"), ("triplet_loss", "models/triplet_loss.py"),
                         ("lightning_model", "models/lightning_model.py"), ("neighbors", "data/utils/neighbors.py"),
                         ("sampler", "data/partition/sampler.py")):
            name = f"{_PKG}." + rel[:-3].replace("/", ".")
            mods[key] = _exec(name, REF_ROOT / rel)
            added.append(name)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in added:
            sys.modules.pop(k, None)
    _cache = SimpleNamespace(fields=fields, geometry_ref=geometry_ref, **mods)
    return _cache


class Frame:
    """The two operations `setup_transcripts_graph` performs on its polars frame
    (`tx[[x_col, y_col]].to_numpy()`, neighbors.py:174), on a dict of numpy columns."""

    def __init__(self, columns):
        self.columns = dict(columns)

    def __getitem__(self, names):
        import numpy as np
        cols = [self.columns[n] for n in names]
        return SimpleNamespace(to_numpy=lambda: np.stack(cols, axis=1))
