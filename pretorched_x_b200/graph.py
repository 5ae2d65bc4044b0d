"""CUDA-graph replay of a whole forward pass.

A forward of resnet3d50 is ~75 kernel launches, each preceded by a Python -> ctypes crossing and a host-side
CUtensorMap encode; at B200 speeds that host work is comparable to the device time.  Capturing the launch
sequence once (static shapes, buffers from the graph's private pool) removes it: the timed step is one
``cudaGraphLaunch``.  All C-ABI entry points are capture-safe (no syncs, no allocations).
"""
import torch


class GraphedForward:
    """``g = GraphedForward(model, example); y = g(x)`` with ``x`` of ``example``'s shape (device or pinned host)."""

    def __init__(self, model, example_input, warmup=2):
        if not example_input.is_cuda:
            raise RuntimeError("GraphedForward needs a CUDA example input: the engine has no CPU path")
        self.model = model
        self.static_in = example_input.detach().clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):          # packs weights, sets kernel attributes, warms the allocator
                model(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = model(self.static_in)

    def __call__(self, x=None):
        if x is not None and x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
