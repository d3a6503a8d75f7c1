"""Host logic of the multi-stream UNet++ inference forward (engine.unetpp_chain_schedule): pure Python, runs on the CPU."""
import pytest

from tactile_gan_b200.engine import unetpp_chain_schedule


@pytest.mark.parametrize("depth,n_streams", [(5, 3), (5, 2), (4, 3), (6, 3), (5, 1)])
def test_unetpp_chain_schedule_respects_every_dependency(depth, n_streams):
    order, stream_of, waits, deps, level = unetpp_chain_schedule(depth, n_streams)
    assert len(order) == depth * (depth + 1) // 2 and len(set(order)) == len(order)
    pos = {nd: k for k, nd in enumerate(order)}
    for nd in order:
        i, j = nd
        # the reference's connectivity (UNet_plusplus.py:72-84): row sources, the upsampled node below, the pooled node above
        want = [(i, k) for k in range(j)] + ([(i + 1, j - 1)] if j else ([(i - 1, 0)] if i else []))
        assert sorted(deps[nd]) == sorted(want)
        for d in deps[nd]:
            assert pos[d] < pos[nd]                       # queued (and its event recorded) before the consumer
            if stream_of[d] == stream_of[nd]:
                assert d not in waits[nd]                 # stream order is enough
            else:
                assert d in waits[nd]                     # cross-stream: an event wait
        assert set(waits[nd]) <= set(deps[nd])
    # every stream sees its nodes by non-decreasing level (no node queued behind a later one of the same stream)
    for s in range(n_streams):
        lv = [level[nd] for nd in order if stream_of[nd] == s]
        assert lv == sorted(lv)


def test_unetpp_chain_schedule_keeps_independent_nodes_apart():
    """Depth 5 on three streams: nodes of the same level (ready at the same time) never share a stream, the critical
    path is 9 of the 15 nodes, and a chain stays on one stream."""
    order, stream_of, waits, deps, level = unetpp_chain_schedule(5, 3)
    by_level = {}
    for nd in order:
        by_level.setdefault(level[nd], []).append(nd)
    assert max(level.values()) + 1 == 9
    for nodes in by_level.values():
        assert len({stream_of[nd] for nd in nodes}) == len(nodes)
    for (i, j) in order:
        if j:
            assert stream_of[i, j] == stream_of[i + 1, j - 1]
