"""CUDA-graph replay of the fake-quant inference forward.

At the reference's evaluation batch (32 images) a calibrated ViT forward is ~600 small launches and the GPU waits for
the host; capturing the whole forward once and replaying it removes the launch latency.  The kernels behind
libadalog_b200.so take the stream they are given and neither allocate nor synchronise, so the forward is capturable as
is (torch's graph-private memory pool keeps every intermediate at a fixed address, which is what the TMA descriptors
of the tensor-core forward need)."""
import torch


class GraphedForward:
    """model(x) for a fixed input shape, captured once and replayed.  `model` must be calibrated and in eval mode."""

    def __init__(self, model, example, warmup=3):
        self.model = model
        self.static_in = example.clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                    # lazy initialisation (cuBLAS handles, caches) outside the capture
                model(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = model(self.static_in)

    def __call__(self, x):
        self.static_in.copy_(x)
        self.graph.replay()
        return self.static_out
