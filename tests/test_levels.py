"""Level table sanity (round-1 advice): a higher level must not compress worse than a lower one on
the corpus -- users who ask for -9 expect at most the size of the default."""
import pytest

from support import KIND_NAMES


@pytest.mark.parametrize("kind", [0, 1], ids=lambda k: KIND_NAMES[k])
def test_sizes_do_not_grow_with_the_level(lib, corpus, kind):
    d = corpus.fill(kind, 416 << 10, offset=5 << 20)
    size = {lvl: len(lib.deflate_bytes(d, lvl)) for lvl in (1, 3, 6, 7, 9)}
    assert size[9] <= size[7] <= size[6] <= size[3] <= size[1], size
