"""CUDA-graph replay of a whole ICL training step (forward, losses, backward, optimizer).

The eager step enqueues ~1 300 kernel launches through Python; at ~25 us of host time per launch the host, not the GPU,
bounds the step (measured: host enqueue 31.6 ms vs 31.8 ms step).  Capturing the step once and replaying it removes that
bound.  Everything on the path is capture-safe: kernels launch on the current (capturing) stream, dropout seeds and the
learning rate live in device memory, DropPath draws use torch's graph-safe CUDA generator, and the optimizer's pointer
table is uploaded from pinned memory.

    step = GraphedStep(lambda x, y: train_step(x, y), (x_example, y_example), optimizer)
    loss = step(x, y)          # copies x, y into the static buffers, refreshes the lr, replays; returns the static loss
"""
import os

import torch


class GraphedStep:
    def __init__(self, fn, example_inputs, optimizer=None, warmup=3):
        self.fn, self.optimizer = fn, optimizer
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # lazily-created state (momentum buffers, function attributes, autograd accumulators)
                self.fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # the step is captured on a high-priority stream: its kernels (the backbone chain) are scheduled ahead of the kernels of the
        # side lanes forked from it (icl_b200/lanes.py), which run at the default priority
        prio = int(os.environ.get("ICL_GRAPH_PRIORITY", "-1"))
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(priority=prio)):
            self.static_out = self.fn(*self.static_in)
        torch.cuda.synchronize()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self.optimizer is not None and hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()
        self.graph.replay()
        return self.static_out

    def release(self):
        """Drop the captured graph (and the NCCL work it holds).  Call before torch.distributed.destroy_process_group():
        tearing the communicator down while a live graph still references its collectives hangs."""
        torch.cuda.synchronize()
        if self.graph is not None:
            self.graph.reset()
        self.graph = None
        self.static_out = None
        torch.cuda.synchronize()
