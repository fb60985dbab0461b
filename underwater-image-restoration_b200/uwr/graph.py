"""CUDA-graph capture of a fixed-shape forward (inference): the ~470 kernel launches of one AST
forward are replayed from a single graph launch, which removes the host launch overhead that
dominates at batch 1 (the eager reference is launch/sync bound at small batch)."""
import torch


class GraphedForward:
    def __init__(self, model, example, warmup=2):
        self.model = model
        self.static_in = example.clone()
        with torch.no_grad():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):     # warm allocator pools, cudaFuncSetAttribute, rounded-weight caches
                    model(self.static_in)
            torch.cuda.current_stream().wait_stream(s)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = model(self.static_in)

    def __call__(self, x):
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedTrainStep:
    """One whole training step (zero -> forward -> loss -> backward -> clip + Adam) captured in a CUDA
    graph and replayed: ~700 (AST) to ~3000 (SpectralTransformer) kernel launches become one.  Everything
    the step touches is graph-safe by construction: gradient buckets and optimizer pointer tables are
    static, the Adam step counter, the learning rate and the clip coefficient live on the device, DropPath masks come from
    torch's graph-aware CUDA generator, TF32-rounded weight copies are refreshed by kernels inside the
    graph.  The `warmup` eager steps are real optimizer steps.

    Data parallel (world > 1) with a raw NCCL communicator (uwr.nccl.Communicator, the default of bench.py): the
    stream-ordered ncclAllReduce calls are captured too — ONE graph, reduces forked onto a side stream.
    Data parallel through torch.distributed work objects: those collectives are kept OUT of the graphs (capturing
    ProcessGroupNCCL work hung on this stack).  The step becomes graph A (zero -> forward -> loss -> backward, hooks only count) ->
    eager per-bucket all-reduce (79.7 MB, ~0.3 ms over NVSwitch, not overlapped) -> graph B (clip + Adam):
    the ~2.5 ms of host launch gaps it removes outweigh the lost overlap."""

    def __init__(self, step, raw, ref, warmup=3):
        self.step = step
        self.raw, self.ref = raw.clone(), ref.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                step(self.raw, self.ref)
        torch.cuda.current_stream().wait_stream(s)
        # the warm-up's activations sit in the caching allocator's default pool; the capture allocates the same amount
        # again in the graph's private pool.  Returning the cached blocks first halves the peak (batch 64 per GPU, BASELINE
        # config 4 at 2 GPUs, needs it).
        # (the last warm-up step's autograd graph is cyclic garbage until Python's collector runs: collect it first)
        import gc
        torch.cuda.synchronize()
        gc.collect()
        torch.cuda.empty_cache()
        self.graph = torch.cuda.CUDAGraph()
        self.graph_opt = None
        if step.world == 1 or step.buckets.comm is not None:
            # one graph for the whole step; with a raw NCCL communicator (uwr.nccl) the bucket all-reduces are
            # captured too, forked onto a side stream as soon as a bucket is complete and joined before clip + Adam
            with torch.cuda.graph(self.graph):
                self.loss, self.norm = step(self.raw, self.ref)
        else:
            step.buckets.overlap = False
            with torch.cuda.graph(self.graph):
                self.loss = step.forward_backward(self.raw, self.ref)
            step.buckets.finish()                      # eager all-reduce of the captured step's buckets
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                self.norm = step.opt.step()

    def replay(self):
        self.step.opt.sync_hyper()     # learning-rate changes (schedulers) reach the captured Adam kernel
        self.graph.replay()
        if self.graph_opt is not None:
            b = self.step.buckets
            b._launched = [False] * len(b.buckets)
            b.finish()
            self.graph_opt.replay()

    def __call__(self, raw, ref):
        self.raw.copy_(raw, non_blocking=True)
        self.ref.copy_(ref, non_blocking=True)
        self.replay()
        return self.loss, self.norm

    # ---- input pipeline: the host -> device copy of step i+1 runs on a side stream under step i ------------
    def prefetch(self, raw_host, ref_host):
        """Start copying the NEXT step's (pinned) host batch into device staging buffers on a copy stream."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._stage = (torch.empty_like(self.raw), torch.empty_like(self.ref))
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        self._copy_stream.wait_event(self._consumed)        # the previous batch has left the staging buffers
        with torch.cuda.stream(self._copy_stream):
            self._stage[0].copy_(raw_host, non_blocking=True)
            self._stage[1].copy_(ref_host, non_blocking=True)
            self._staged.record()

    def step_prefetched(self):
        """Run one step on the batch handed to prefetch(): wait for its copy, move it into the graph's static
        inputs (device -> device, microseconds) and replay."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        self.raw.copy_(self._stage[0], non_blocking=True)
        self.ref.copy_(self._stage[1], non_blocking=True)
        self._consumed.record(cur)
        self.replay()
        return self.loss, self.norm
