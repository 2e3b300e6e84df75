"""Host-side logic that needs no GPU."""
import pytest


@pytest.mark.parametrize("nq", [1, 1023, 1024, 2048, 2049, 4096, 5000, 8192, 20000, 65536])
def test_pipe_chunks_tile_the_batch_and_shrink_towards_the_tail(nq):
    from movie_recommender_demo_b200.faiss_retrieval import _PIPE_CHUNK, _pipe_chunks
    ch = _pipe_chunks(nq)
    assert ch[0][0] == 0 and ch[-1][1] == nq
    assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
    sizes = [hi - lo for lo, hi in ch]
    assert all(s > 0 for s in sizes)
    assert sizes[-1] <= _PIPE_CHUNK                       # the un-hidden tail copy stays small
    assert all(a <= 4 * b for a, b in zip(sizes, sizes[1:]))   # each copy hides under the next chunk's search


def test_dimension_is_validated_before_any_gpu_work():
    """Constructor argument checks come first, so they also fire on a CPU-only box (no silent fallback:
    a valid dimension then fails on the missing GPU)."""
    from movie_recommender_demo_b200 import _lib
    from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
    FAISSIndex.verbose = False
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises((RuntimeError, _lib.B2RError, OSError)):
        FAISSIndex(256, 'Flat')
    with pytest.raises(ValueError, match="Unknown index type"):
        FAISSIndex(256, 'Nope')
