"""Host read-back pipeline (dist.HostFrameSink): copies issued behind later kernels land complete and in order."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_host_frame_sink_double_buffer():
    from image2video_synthesis_using_cinns_b200.dist import HostFrameSink
    sink = HostFrameSink("cuda")
    assert sink.wait() is None
    outs = []
    for k in range(5):
        x = torch.full((3, 4, 3, 8, 8), float(k), device="cuda") + torch.arange(8, device="cuda")
        buf = sink.put(x)
        torch.zeros(1 << 20, device="cuda").normal_()      # unrelated later work on the producer's stream overlaps the copy
        outs.append((k, buf))
        got = sink.wait()
        assert got is buf
        assert torch.equal(got, torch.full((3, 4, 3, 8, 8), float(k)) + torch.arange(8.0))
    # two alternating pinned buffers
    assert outs[0][1] is outs[2][1] and outs[1][1] is outs[3][1] and outs[0][1] is not outs[1][1]
    assert outs[0][1].is_pinned()
