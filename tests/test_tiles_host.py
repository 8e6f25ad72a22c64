"""Host logic of segger_b200.tiles (no GPU): the bin packing against the reference's own functions."""
import random

import pytest

from oracle import reference_import
from segger_b200 import tiles

live = pytest.mark.skipif(not reference_import.available(), reason="reference tree only exists in the build container")


def test_best_fit_decreasing_properties():
    rng = random.Random(0)
    items = [rng.randint(1, 700) for _ in range(400)]
    bins = tiles.best_fit_decreasing(items, 1000)
    assert sorted(i for b in bins for i in b) == list(range(400))
    assert all(sum(items[i] for i in b) <= 1000 for b in bins)
    assert len(bins) <= 1.25 * sum(items) / 1000 + 1                   # BFD is within 11/9 OPT + 1
    with pytest.raises(ValueError):
        tiles.best_fit_decreasing([5, 0], 10)
    with pytest.raises(ValueError):
        tiles.best_fit_decreasing([5, 11], 10)
    assert tiles.best_fit_decreasing([5, 0, 11, 4], 10, skip_too_big=True) == [[0, 3]]
    assert tiles.best_fit_decreasing([], 10) == []


@live
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_packing_equals_reference_sampler(seed):
    """best_fit_decreasing / first_fit_decreasing_bucketed of data/partition/sampler.py:11-82,186-282, run as is."""
    ref = reference_import.load().sampler
    rng = random.Random(seed)
    items = [rng.randint(1, 900) for _ in range(500)] + [1000, 1000, 1]
    assert tiles.best_fit_decreasing(items, 1000) == ref.best_fit_decreasing(items, 1000)
    assert (tiles.first_fit_shuffled(items, 1000, rng=random.Random(seed + 10))
            == ref.first_fit_decreasing_bucketed(items, 1000, rng=random.Random(seed + 10)))
    more = items + [0, -3, 4000]
    assert (tiles.best_fit_decreasing(more, 1000, skip_too_big=True)
            == ref.best_fit_decreasing(more, 1000, skip_too_big=True))
    fl = [rng.uniform(0.01, 1.0) for _ in range(200)]                   # float weights, ties unlikely
    assert tiles.best_fit_decreasing(fl, 1.0) == ref.best_fit_decreasing(fl, 1.0)
    ties = [rng.choice([100, 200, 300]) for _ in range(300)]            # many equal sizes: tie rule = lowest bin index
    assert tiles.best_fit_decreasing(ties, 1000) == ref.best_fit_decreasing(ties, 1000)
