"""CPU: host logic of graphs.GraphCache (one captured graph per input-shape key, LRU eviction, optional padding of the spaced
text to width buckets) with the capture itself replaced by a recorder."""
import torch

from handwriting_line_generation_b200 import graphs


def test_graph_cache_captures_once_per_shape_and_evicts_lru(monkeypatch):
    made = []

    class FakeGraphed:
        def __init__(self, fn, example_inputs, modules=(), warmup=3):
            self.fn, self.shapes = fn, [tuple(x.shape) for x in example_inputs]
            made.append(self.shapes)

        def __call__(self, *inputs):
            assert [tuple(x.shape) for x in inputs] == self.shapes
            return self.fn(*inputs)

    monkeypatch.setattr(graphs, "GraphedStep", FakeGraphed)
    step = graphs.GraphCache(lambda c, s: c.sum() + s.sum(), max_graphs=2)
    s = torch.ones(2, 4)
    for T in (8, 8, 12, 8, 16, 12):
        assert float(step(torch.ones(T, 2, 5), s)) == T * 10 + 8
    # 8 captured, replayed; 12 captured; 8 replayed (now most recent); 16 evicts 12; 12 captured again
    assert [sh[0][0] for sh in made] == [8, 12, 16, 12] and step.captures == 4
    padded = graphs.GraphCache(lambda c, s: c, max_graphs=4, bucket=graphs.pad_spaced_text(32))
    out = padded(torch.nn.functional.one_hot(torch.randint(1, 5, (40, 2)), 5).float(), s)
    assert tuple(out.shape) == (64, 2, 5) and float(out[40:, :, 0].min()) == 1.0 and float(out[40:, :, 1:].max()) == 0.0
    padded(torch.zeros(64, 2, 5), s)
    assert padded.captures == 1
