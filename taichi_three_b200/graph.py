"""Frames as CUDA graphs.

Small scenes (C1: 968 faces, C4: 34 faces x 3 objects x 64 views) spend their time between kernels, not in
them: a `Scene.render()` of three objects is ~17 launches issued from Python.  `FrameGraph(fn)` records whatever
`fn` enqueues -- `scene.render()`, a loop over cameras, raw engine/raster calls -- once, and `replay()` submits
the whole frame with one driver call.

What a recorded frame freezes: everything that travels as a kernel argument at record time -- camera matrices,
sample bias, materials, lights, face counts, buffer addresses.  What it re-reads at every replay: the contents
of device buffers (mesh positions / normals, textures, the image).  So animate geometry in place
(`mesh.pos.to_torch()[...] = ...`) and replay; re-record after changing the camera, a material or a light.

The C ABI detects stream capture by itself (`cudaStreamIsCapturing`): under capture `render_occup` always
records the tile-path kernel and zeroes its own counter set, and nothing is published to the host.  Inputs must be
device resident (host arrays would be baked in as pageable-memory copies, which capture refuses) and every buffer
must have reached its final size, hence the eager warm-up calls before recording.
"""
import torch


class FrameGraph:
    def __init__(self, fn, warmup=2):
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            fn()

    def replay(self):
        self.graph.replay()

    __call__ = replay
